cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "bucketed or resident_index or super_chunks" 2>&1 | tail -5
q() {  # label, config, env...
  local label="$1" cfg="$2" scale="$3"; shift 3
  env "$@" timeout 900 python bench.py --config $cfg --scale $scale --steps 5 --warmup 3 --no-cpu --no-e2e --no-target 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$label cfg$cfg x$scale', 'warm ms', round(d['ms_per_step'],3), 'score', round(d['phase_ms_rank0']['score'],3), 'plain score', round(d['roofline']['plain_index']['score_ms'],3), 'build', round(d['index_build_ms_rank0'],2), 'cold ms', round(d['ms_per_step_cold'],3), 'extract', round(d['phase_ms_rank0']['extract_cold'],3), 'hits', d['hits_per_step'], 'frac', round(d['roofline']['frac'],3), 'plain frac', round(d['roofline']['plain_index']['frac'],3))
except Exception as e: print('$label cfg$cfg failed', e)"
}
q base 3 0.25 A=1
q base 4 0.25 A=1
q base 3 1.0 A=1
q base 4 1.0 A=1
q key6 3 0.25 VARSCOT_LIB=/root/repo/build/variants/lib_key6.so; q key6 4 0.25 VARSCOT_LIB=/root/repo/build/variants/lib_key6.so; for c in 32; do q ctas$c 3 0.25 VARSCOT_SCORE_CTAS_PER_SM=$c; q ctas$c 4 0.25 VARSCOT_SCORE_CTAS_PER_SM=$c; done
CMD="python bench.py --config 4 --scale 0.25 --steps 2 --warmup 3 --no-cpu --no-e2e --no-target"
ncu --set full --clock-control none --import-source on -k regex:k_score_bucketed -s 2 -c 1 -o gpurun_out/r2_bk_cfg4 $CMD > gpurun_out/ncu_bk4.log 2>&1
echo "bk cfg4 rc=$?"
