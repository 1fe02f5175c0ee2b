for cg in 0 128 256; do
  DM_RA_CG=$cg timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/s12.json 2> gpurun_out/s12.err
  python -c "
import json; d=json.load(open('gpurun_out/s12.json')); print('cg $cg', {k:round(v['ms'],3) for k,v in d['kernels'].items()}, d['checksums'])" 2>&1 | tail -1
done
