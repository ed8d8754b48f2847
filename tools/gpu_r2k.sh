# round 2, 2-GPU session K: validates the multi-rank paths of bench.py at a reduced scale before the 8-GPU session
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --scale 0.25 --steps 5 --warmup 3 \
    > gpurun_out/r2k_cfg3_2gpu.json 2> gpurun_out/r2k_cfg3_2gpu.err; echo "N=2 rc=$?"; tail -3 gpurun_out/r2k_cfg3_2gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2k_cfg3_2gpu.json').read().strip().splitlines()[-1])
    print('N=2 value',round(d['value']),'ms',round(d['ms_per_step'],3),'plain',round(d['value_plain_index']),'cold',round(d['value_cold']),'scaling',d['scaling'],'hits',d['hits_per_step'])
    print('  e2e',d['e2e'] and (round(d['e2e']['value']),round(d['e2e']['ms_per_step'],2)),'rg',d['e2e_resident_genome'],'parity',d.get('parity'))
    t=d.get('target_cfg4'); print('  cfg4',t and (round(t['value']),round(t['ms_per_step'],2),t.get('parity'),t.get('e2e') and round(t['e2e']['ms_per_step'],2)))
except Exception as e: print('parse failed',e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29703 bench.py --gpus 2 --config 5 --scale 0.01 --guides 2000 \
    > gpurun_out/r2k_cfg5_2gpu.json 2> gpurun_out/r2k_cfg5_2gpu.err; echo "cfg5 rc=$?"; tail -3 gpurun_out/r2k_cfg5_2gpu.err; tail -c 1200 gpurun_out/r2k_cfg5_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29704 bench.py --gpus 2 --impl reference --config 1 --steps 1 --warmup 1 | tail -c 400
build/h2d8 2 | tail -c 900
