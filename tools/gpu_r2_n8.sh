# round 2, 8-GPU session: H2D ceiling, the strong-partition bench at N = 2, 4, 8 (config 3 + config-4 target block), config 5 at full size
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
( lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" ; cat /sys/devices/system/node/node*/meminfo 2>/dev/null | grep MemTotal ) > gpurun_out/r2_host.txt 2>&1
build/h2d8 > gpurun_out/r2_h2d8.json 2> gpurun_out/r2_h2d8.err; echo "h2d8 rc=$?"; cat gpurun_out/r2_h2d8.json
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 10 --warmup 3 ${NOCPU:-} \
      > gpurun_out/r2_bench_cfg3_${n}gpu.json 2> gpurun_out/r2_bench_cfg3_${n}gpu.err
  echo "N=$n rc=$?"; tail -2 gpurun_out/r2_bench_cfg3_${n}gpu.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_cfg3_${n}gpu.json').read().strip().splitlines()[-1])
    print('N=$n value',round(d['value']),'ms',round(d['ms_per_step'],3),'plain',round(d['value_plain_index']),'cold',round(d['value_cold']),round(d['ms_per_step_cold'],3))
    print('  e2e',d['e2e'] and (round(d['e2e']['value']),round(d['e2e']['ms_per_step'],2)),'rg',d['e2e_resident_genome'] and (round(d['e2e_resident_genome']['value']),round(d['e2e_resident_genome']['ms_per_step'],2)),'parity',d.get('parity',{}).get('diff'))
    t=d.get('target_cfg4'); print('  cfg4',t and (round(t['value']),round(t['ms_per_step'],2),round(t['frac_executed'],3),'plain',round(t['plain_index']['ms_per_step'],2),t.get('parity',{}).get('diff'),t.get('e2e') and round(t['e2e']['ms_per_step'],2)))
except Exception as e: print('parse failed',e)
PY
done
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --config 5 \
    > gpurun_out/r2_bench_cfg5_8gpu.json 2> gpurun_out/r2_bench_cfg5_8gpu.err
echo "cfg5 rc=$?"; tail -3 gpurun_out/r2_bench_cfg5_8gpu.err; tail -c 1500 gpurun_out/r2_bench_cfg5_8gpu.json
