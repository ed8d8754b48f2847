# round 2, GPU session A: parity tests, the bench line of config 3, then quick timings of the k_score tuning builds
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -40
echo "== bench cfg3"
timeout 900 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err; echo "rc=$?"; tail -3 gpurun_out/r2a_bench_cfg3.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2a_bench_cfg3.json').read().strip().splitlines()[-1])
    print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'plain',round(d['value_plain_index']),round(d['ms_per_step_plain_index'],3),'build',round(d['index_build_ms_rank0'],2),'cold',round(d['value_cold']),round(d['ms_per_step_cold'],3),'phase',d['phase_ms_rank0'])
    print('e2e',d['e2e'] and (round(d['e2e']['value']),round(d['e2e']['ms_per_step'],2)),'rg',d['e2e_resident_genome'] and (round(d['e2e_resident_genome']['value']),round(d['e2e_resident_genome']['ms_per_step'],2)))
    print('roof frac',round(d['roofline']['frac'],3),'lds',round(d['roofline']['frac_lds'],3),'plain frac',round(d['roofline']['plain_index']['frac'],3),'parity',d.get('parity'),'cpu',d.get('cpu_baseline',{}).get('value'))
    t=d.get('target_cfg4'); print('cfg4',t and (round(t['value']),round(t['ms_per_step'],2),round(t['frac_executed'],3),'plain',round(t['plain_index']['ms_per_step'],2),round(t['plain_index']['frac_executed'],3),'build',round(t['index_build_ms_rank0'],1),t.get('parity'),t.get('e2e') and round(t['e2e']['ms_per_step'],2)))
except Exception as e: print('parse failed',e)
PY
q() {  # label, config, env...
  local label="$1" cfg="$2"; shift 2
  env "$@" timeout 600 python bench.py --config $cfg --scale 0.25 --steps 5 --warmup 3 --no-cpu --no-e2e --no-target 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$label cfg$cfg', 'warm ms', round(d['ms_per_step'],3), 'score', round(d['phase_ms_rank0']['score'],3), 'plain score', round(d['roofline']['plain_index']['score_ms'],3), 'build', round(d['index_build_ms_rank0'],2), 'cold ms', round(d['ms_per_step_cold'],3), 'extract', round(d['phase_ms_rank0']['extract_cold'],3), 'hits', d['hits_per_step'], 'frac', round(d['roofline']['frac'],3))
except Exception as e: print('$label cfg$cfg failed', e)"
}
q base 3 A=1
q base 4 A=1
for v in w8; do
  q $v 3 VARSCOT_LIB=/root/repo/build/variants/lib_$v.so
  q $v 4 VARSCOT_LIB=/root/repo/build/variants/lib_$v.so
done
for c in 32; do q ctas$c 3 VARSCOT_SCORE_CTAS_PER_SM=$c; q ctas$c 4 VARSCOT_SCORE_CTAS_PER_SM=$c; done
