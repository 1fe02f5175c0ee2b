# Host <-> device copy ceiling and the bench line at N ranks of one box.
# usage (under gpurun --gpus N): bash tools/gpu/r02_pcie.sh <tag> <N>
TAG=$1; N=$2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
nvidia-smi topo -m > gpurun_out/${TAG}_topo_$N.log 2>&1
lscpu | grep -i -E "numa|socket|model name|^cpu\(s\)" > gpurun_out/${TAG}_lscpu_$N.log 2>&1
timeout 600 $RUN tools/pcie_ceiling.py > gpurun_out/${TAG}_pcie_$N.json 2> gpurun_out/${TAG}_pcie_$N.err; echo "pcie rc=$?"
cat gpurun_out/${TAG}_pcie_$N.json
timeout 900 $RUN bench.py --gpus $N --no-competitor > gpurun_out/${TAG}_bench_$N.json 2> gpurun_out/${TAG}_bench_$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d.get('e2e'))
print({k:(v.get('img_per_s') if isinstance(v,dict) else None) for k,v in d.get('configs',{}).items()})
PY
