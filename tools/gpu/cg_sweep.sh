# channels per work unit of the multi-bucket launch (DM_RA_CG overrides every bucket)
for cg in 0 128 0 128; do
  DM_RA_CG=$cg timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/cg.json 2> gpurun_out/cg.err
  python -c "
import json; d=json.load(open('gpurun_out/cg.json')); print('cg $cg', {k:round(v['ms'],3) for k,v in d['kernels'].items()})" 2>&1 | tail -1
done
