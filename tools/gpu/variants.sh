# usage: variants.sh name1 name2 ...  -- bench fwd/bwd times for tools/bin/lib_<name>.so builds ("main" = the product library)
for v in "$@"; do
  if [ "$v" = main ]; then unset DYNAMASK_LIB; else export DYNAMASK_LIB=$PWD/tools/bin/lib_$v.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/var_$v.json')); print('$v', round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['kernels'].items()}, d['checksums'])" 2>&1 | tail -1
done
