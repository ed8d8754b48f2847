# round 2, GPU session J: the whole GPU suite, smoke(), the full bench line (+ --cli leg) and the reference arm
cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -8
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1200 python bench.py --config 3 --steps 10 --warmup 3 --cli > gpurun_out/r2_bench_cfg3_1gpu.json 2> gpurun_out/r2_bench_cfg3_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_cfg3_1gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_cfg3_1gpu.json').read().strip().splitlines()[-1])
    print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'plain',round(d['value_plain_index']),round(d['ms_per_step_plain_index'],3),'build',round(d['index_build_ms_rank0'],2),'cold',round(d['value_cold']),round(d['ms_per_step_cold'],3),'phase',d['phase_ms_rank0'])
    print('e2e',d['e2e'] and (round(d['e2e']['value']),round(d['e2e']['ms_per_step'],2)),'rg',d['e2e_resident_genome'] and (round(d['e2e_resident_genome']['value']),round(d['e2e_resident_genome']['ms_per_step'],2)))
    print('roof frac',round(d['roofline']['frac'],3),'lds',round(d['roofline']['frac_lds'],3),'plain frac',round(d['roofline']['plain_index']['frac'],3),'parity',d.get('parity',{}).get('diff'),'cpu',d.get('cpu_baseline',{}).get('value'))
    t=d.get('target_cfg4'); print('cfg4',t and (round(t['value']),round(t['ms_per_step'],2),round(t['frac_executed'],3),'plain',round(t['plain_index']['ms_per_step'],2),round(t['plain_index']['frac_executed'],3),'build',round(t['index_build_ms_rank0'],1),t.get('parity'),t.get('e2e') and round(t['e2e']['ms_per_step'],2)))
    print('cli',d.get('cli_e2e'))
except Exception as e: print('parse failed',e)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "ref rc=$?"; tail -c 700 gpurun_out/r2_bench_reference_arm.json
for c in 1 2; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg${c}_1gpu.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_cfg${c}_1gpu.json').read().strip().splitlines()[-1])
print('cfg$c value',round(d['value']),'ms',round(d['ms_per_step'],3),'plain',round(d['value_plain_index']),'cold',round(d['value_cold']),round(d['ms_per_step_cold'],3),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],2),'parity',d.get('parity',{}).get('diff'),'cpu',round(d['cpu_baseline']['value'],1))"; done
