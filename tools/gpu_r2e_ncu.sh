cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --config 4 --scale 0.25 --steps 2 --warmup 3 --no-cpu --no-e2e --no-target"
ncu --set full --clock-control none --import-source on -k regex:k_score_bucketed -s 2 -c 1 -o gpurun_out/r2_bk_cfg4 $CMD > gpurun_out/ncu_bk4.log 2>&1
echo "bk cfg4 rc=$?"
CMD="python bench.py --config 3 --scale 0.25 --steps 2 --warmup 3 --no-cpu --no-e2e --no-target"
ncu --set full --clock-control none --import-source on -k regex:k_score_bucketed -s 2 -c 1 -o gpurun_out/r2_bk_cfg3 $CMD > gpurun_out/ncu_bk3.log 2>&1
echo "bk cfg3 rc=$?"
