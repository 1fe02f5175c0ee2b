mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/scale4.json 2> gpurun_out/scale4.err; echo rc=$?
tail -3 gpurun_out/scale4.err
python - <<'PY'
import json
for l in open('gpurun_out/scale4.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','checksums','gpu_launches')}); print('e2e',d.get('e2e'))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 1 --warmup 1 --impl reference 2>/dev/null | cut -c1-300
