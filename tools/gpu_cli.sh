cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests -x -q -m gpu -k "multi_device or cli" 2>&1 | tail -4
python tools/cli_bench.py 600000000 800000 100 6 2>&1 | tail -3 | tee gpurun_out/cli_bench.txt
