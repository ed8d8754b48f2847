cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2 | tail -1
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "n$n rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('n_gpus','value','ms_per_step','wall_ms_per_step','phase_ms')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
