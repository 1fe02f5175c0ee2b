# RED + zero-fill prefetch check: parity tests, bench, ring-depth variant
bash tools/gpu/quick.sh s9
for kb in 112; do
  DYNAMASK_LIB=$PWD/tools/bin/lib_ring16.so DM_RA_BWD_SMEM_KB=$kb timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/s9_ring16.json 2> gpurun_out/s9_ring16.err
  python -c "
import json; d=json.load(open('gpurun_out/s9_ring16.json')); print('ring16 smem $kb', {k:round(v['ms'],3) for k,v in d['kernels'].items()}, d['checksums'])" 2>&1 | tail -1
done
