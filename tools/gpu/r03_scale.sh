# bench.py at N ranks exactly as the driver launches it, both arms
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r03_scale$N.json 2> gpurun_out/r03_scale$N.err; echo rc=$?
tail -3 gpurun_out/r03_scale$N.err
python - <<PY
import json
for l in open('gpurun_out/r03_scale$N.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','gpu_launches')}); print('e2e',d.get('e2e'))
        for k,v in d.get('configs',{}).items(): print(k, {kk:vv for kk,vv in v.items() if kk in ('img_per_s','images','ms_per_pass','scaling')} if 'img_per_s' in v else {kk:vv.get('img_per_s') for kk,vv in v.items() if isinstance(vv,dict) and 'img_per_s' in vv})
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1 --warmup 1 --impl reference 2> gpurun_out/r03_scale${N}_ref.err | cut -c1-400; echo refrc=$?
