cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --scale 0.25 --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_score -s 2 -c 2 -o gpurun_out/r1_score $CMD > gpurun_out/ncu_score.log 2>&1
echo "score rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_extract -s 1 -c 1 -o gpurun_out/r1_extract $CMD > gpurun_out/ncu_extract.log 2>&1
echo "extract rc=$?"
tail -3 gpurun_out/ncu_score.log
ls -la gpurun_out
