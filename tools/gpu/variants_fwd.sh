# forward occupancy variants x shared-memory size
for v in main f256r128 f384r80 f256r80; do
  for kb in 72 100; do
    if [ "$v" = main ]; then unset DYNAMASK_LIB; else export DYNAMASK_LIB=$PWD/tools/bin/lib_$v.so; fi
    DM_RA_FWD_SMEM_KB=$kb timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/varf.json 2> gpurun_out/varf.err
    python -c "
import json; d=json.load(open('gpurun_out/varf.json')); print('$v smem $kb', {k:round(v['ms'],3) for k,v in d['kernels'].items()}, d['checksums'])" 2>&1 | tail -1
  done
done
