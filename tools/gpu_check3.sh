# parity tests + one full bench run (config 3) with a short summary
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
print({k:d[k] for k in ('value','ms_per_step','wall_ms_per_step','phase_ms','hits_per_step','gpu_launches')})
print('e2e',d['e2e']); print('roof',{k:d['roofline'][k] for k in ('frac','frac_executed','frac_lds','avg_launch_ms')}); print(d.get('parity'), d.get('cpu_baseline',{}).get('value'), d.get('clocks'))
PY
tail -3 gpurun_out/bench_full.err
