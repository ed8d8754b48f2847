# round 2, GPU session C: ncu evidence for k_score (guide per lane) and k_extract on the full-size config 3 / 4 text
cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --config 3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-target"
CMD4="python bench.py --config 4 --steps 2 --warmup 3 --no-cpu --no-e2e --no-target"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg3.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
# the 17th k_score launch: 14 per-chunk launches of the first (index-building) scan, then whole-store launches of the warm scans
ncu --set full --clock-control none --import-source on -k regex:k_score -s 16 -c 1 -o gpurun_out/r2_score_cfg3 $CMD > gpurun_out/ncu_score3.log 2>&1
echo "score cfg3 rc=$?"
$CMD4 > gpurun_out/plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_score -s 16 -c 1 -o gpurun_out/r2_score_cfg4 $CMD4 > gpurun_out/ncu_score4.log 2>&1
echo "score cfg4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_extract\$ -s 6 -c 1 -o gpurun_out/r2_extract_cfg3 $CMD > gpurun_out/ncu_extract.log 2>&1
echo "extract rc=$?"
ls -la gpurun_out/*.ncu-rep
