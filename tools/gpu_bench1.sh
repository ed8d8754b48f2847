cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2
python bench.py --scale 0.05 --steps 3 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "small rc=$?"; tail -c 3000 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full rc=$?"; tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
