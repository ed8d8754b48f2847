cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --scale 0.25 --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches_final.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_score -s 6 -c 1 -o gpurun_out/r1_score_final $CMD > gpurun_out/ncu_score.log 2>&1
echo "score rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_extract -s 6 -c 1 -o gpurun_out/r1_extract_final $CMD > gpurun_out/ncu_extract.log 2>&1
echo "extract rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3_final.json 2> gpurun_out/bench_cfg3_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"
tail -c 600 gpurun_out/bench_ref_final.json
