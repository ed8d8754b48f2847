cd /root/repo
mkdir -p gpurun_out
run() {
  touch varscot_b200/csrc/vs_device.cu
  make EXTRA="$1" > /dev/null 2>&1 || { echo "build failed: $1"; return; }
  python bench.py --config 4 --scale 0.1 --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms', round(d['ms_per_step'],3), 'extract', round(d['phase_ms']['extract'],3), 'score', round(d['phase_ms']['score'],3), 'hits', d['hits_per_step'], 'launches', d['gpu_launches'])"
}
run "-DVS_SCORE_THREADS=96 -DVS_SCORE_MINBLOCKS=6 -DVS_PAT_CHUNK=320"
run "-DVS_SCORE_THREADS=96 -DVS_SCORE_MINBLOCKS=6 -DVS_PAT_CHUNK=160"
run "-DVS_SCORE_THREADS=96 -DVS_SCORE_MINBLOCKS=6 -DVS_PAT_CHUNK=96"
run "-DVS_SCORE_THREADS=96 -DVS_SCORE_MINBLOCKS=6 -DVS_PAT_CHUNK=64"
run "-DVS_SCORE_THREADS=256 -DVS_SCORE_MINBLOCKS=2 -DVS_PAT_CHUNK=320"
run "-DVS_SCORE_THREADS=256 -DVS_SCORE_MINBLOCKS=2 -DVS_PAT_CHUNK=160"
run "-DVS_SCORE_THREADS=128 -DVS_SCORE_MINBLOCKS=4 -DVS_PAT_CHUNK=256"
