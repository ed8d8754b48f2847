# quick timings of `make EXTRA=...` build variants (config 3 at 0.25 scale; no parity tests: run gpu_exp.sh for those)
cd /root/repo
mkdir -p gpurun_out
q() {
  local label="$1" cfg="$2"
  python bench.py --config $cfg --scale 0.25 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$label cfg$cfg', 'ms', round(d['ms_per_step'],3), 'extract', round(d['phase_ms']['extract'],3), 'score', round(d['phase_ms']['score'],3), 'hits', d['hits_per_step'])"
}
q base 3
for v in "$@"; do
  touch varscot_b200/csrc/vs_device.cu
  make EXTRA="$v" > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  q "[$v]" 3
done
