# experiment batch: parity tests, then quick timings of build variants (configs 3, 2, 4 at 0.25 scale)
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
q() {  # label, config, env...
  local label="$1" cfg="$2"; shift 2
  env "$@" python bench.py --config $cfg --scale 0.25 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$label cfg$cfg', 'ms', round(d['ms_per_step'],3), 'extract', round(d['phase_ms']['extract'],3), 'score', round(d['phase_ms']['score'],3), 'hits', d['hits_per_step'])"
}
q base 3 A=1
q base 2 A=1
q base 4 A=1
for v in "$@"; do
  touch varscot_b200/csrc/vs_device.cu
  make EXTRA="$v" > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  q "[$v]" 3 A=1
  q "[$v]" 2 A=1
  q "[$v]" 4 A=1
done
