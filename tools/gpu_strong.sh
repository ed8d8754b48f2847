cd /root/repo
mkdir -p gpurun_out
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 10 --warmup 3 --scaling strong > gpurun_out/bench_strong_n$n.json 2> gpurun_out/bench_strong_n$n.err; echo "n$n rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_strong_n$n.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('n_gpus','scaling','value','ms_per_step','wall_ms_per_step','phase_ms')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'])
PY
done
