# round 2, GPU session M: A/B of the k_score tuning knobs, then the final ncu evidence (launch list + --set full captures) at full scale
cd /root/repo
mkdir -p gpurun_out
# everything written after the last GPU session first: the whole GPU suite (tests/test_zz_gpu_round2_late.py last), smoke, the default line
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -8
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r2m_bench_cfg3_1gpu.json 2> gpurun_out/r2m_bench_cfg3_1gpu.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2m_bench_cfg3_1gpu.json
q() {  # label, config, scale, env...
  local label="$1" cfg="$2" scale="$3"; shift 3
  env "$@" timeout 900 python bench.py --config $cfg --scale $scale --steps 8 --warmup 3 --no-cpu --no-e2e --no-target 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$label cfg$cfg x$scale', d['index'], 'warm ms', round(d['ms_per_step'],3), 'score', round(d['phase_ms_rank0']['score'],3), 'plain score', round(d['roofline']['plain_index']['score_ms'],3), 'cold ms', round(d['ms_per_step_cold'],3), 'score_cold', round(d['phase_ms_rank0']['score_cold'],3), 'extract', round(d['phase_ms_rank0']['extract_cold'],3), 'plain frac', round(d['roofline']['plain_index']['frac'],3))
except Exception as e: print('$label cfg$cfg failed', e)"
}
q default 3 1.0 A=1
q norot 3 1.0 VARSCOT_SCORE_ROT=40
q nofold 3 1.0 VARSCOT_SCORE_FOLD_TAIL=0
q norot_nofold 3 1.0 VARSCOT_SCORE_ROT=40 VARSCOT_SCORE_FOLD_TAIL=0
q rot0 3 1.0 VARSCOT_SCORE_ROT=0
q ctas16 3 1.0 VARSCOT_SCORE_CTAS_PER_SM=16
CMD="python bench.py --config 3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-target"
CMD4="python bench.py --config 4 --steps 2 --warmup 3 --no-cpu --no-e2e --no-target"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cfg3.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
# k_score: the 17th launch = the first whole-store launch of the plain index (14 per-chunk launches of the index-building scan, 2 warm-ups)
ncu --set full --clock-control none --import-source on -k regex:k_score\$ -s 16 -c 1 -o gpurun_out/r2_score_cfg3 $CMD > gpurun_out/ncu_score3.log 2>&1; echo "score cfg3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_score_bucketed -s 2 -c 1 -o gpurun_out/r2_bucketed_cfg3 $CMD > gpurun_out/ncu_bk3.log 2>&1; echo "bucketed cfg3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_score_bucketed -s 2 -c 1 -o gpurun_out/r2_bucketed_cfg4 $CMD4 > gpurun_out/ncu_bk4.log 2>&1; echo "bucketed cfg4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_score\$ -s 16 -c 1 -o gpurun_out/r2_score_cfg4 $CMD4 > gpurun_out/ncu_score4.log 2>&1; echo "score cfg4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_extract\$ -s 6 -c 1 -o gpurun_out/r2_extract_cfg3 $CMD > gpurun_out/ncu_extract.log 2>&1; echo "extract rc=$?"
ncu --set full --clock-control none -k regex:k_bucket_gather -s 0 -c 1 -o gpurun_out/r2_bucket_gather_cfg3 $CMD > gpurun_out/ncu_gather.log 2>&1; echo "gather rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -8
