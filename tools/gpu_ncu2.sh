cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
CMD="python bench.py --scale 0.25 --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_score -s 6 -c 1 -o gpurun_out/r1_score_v2 $CMD > gpurun_out/ncu_score.log 2>&1
echo "score rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_extract -s 6 -c 1 -o gpurun_out/r1_extract_v3 $CMD > gpurun_out/ncu_extract.log 2>&1
echo "extract rc=$?"
tail -c 1500 gpurun_out/plain3.log
