cd /root/repo
mkdir -p gpurun_out
for c in 2 4 1; do
  python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err; echo "cfg$c rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg$c.json'))
print({k:d[k] for k in ('value','ms_per_step','phase_ms','hits_per_step','gpu_launches')})
print('e2e',{k:d['e2e'][k] for k in ('value','ms_per_step','h2d_bytes_per_step')}); print('roof',{k:d['roofline'][k] for k in ('frac','frac_executed','frac_lds')}); print(d.get('parity'), d.get('cpu_baseline',{}).get('value'))
PY
done
