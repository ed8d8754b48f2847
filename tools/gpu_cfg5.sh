cd /root/repo
mkdir -p gpurun_out
python bench.py --config 5 --scale 0.02 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_cfg5_s002.json 2> gpurun_out/bench_cfg5.err; echo "rc=$?"; tail -3 gpurun_out/bench_cfg5.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg5_s002.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','wall_ms_per_step','phase_ms','hits_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
python bench.py --config 3 --guides 100 --scale 0.1 --steps 3 --warmup 3 > gpurun_out/b.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print(d['value'], d['parity'], d['config'].get('cpus_bound_per_rank'), d['cpu_baseline']['cores'])"
