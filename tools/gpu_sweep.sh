cd /root/repo
mkdir -p gpurun_out
run() {
  touch varscot_b200/csrc/vs_device.cu
  make EXTRA="$1" > /dev/null 2>&1 || { echo "build failed: $1"; return; }
  VARSCOT_TILE_WORDS=$2 python bench.py --scale 0.25 --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 tile=$2', 'ms', round(d['ms_per_step'],3), 'extract', round(d['phase_ms']['extract'],3), 'score', round(d['phase_ms']['score'],3), 'hits', d['hits_per_step'])"
}
run "-DVS_EX_MINBLOCKS=10" 240
run "-DVS_EX_MINBLOCKS=12" 240
run "-DVS_EX_MINBLOCKS=14" 240
run "-DVS_EX_MINBLOCKS=16" 240
run "-DVS_EX_MINBLOCKS=8" 240
run "-DVS_EX_MINBLOCKS=10" 248
run "-DVS_EX_MINBLOCKS=10" 224
run "-DVS_EX_MINBLOCKS=10" 120
run "-DVS_EX_MINBLOCKS=10" 256
