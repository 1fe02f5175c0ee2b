mkdir -p gpurun_out
cp dynamask_b200/lib/libdynamask_sm100.so /tmp/orig.so
run() { cp dynamask_b200/lib/variants/$1.so dynamask_b200/lib/libdynamask_sm100.so; DM_RA_FWD_SMEM_KB=$2 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 smem $2KB: fwd %.3f ms bwd %.3f ms'%(d['kernels']['dm_roi_align_fwd']['ms'], d['kernels']['dm_roi_align_bwd(+zero-init)']['ms']))"; }
run fwd_256_128 110
run diag_NO_STORES 110
run diag_PLAIN_STORES 110
run diag_NO_READS 110
cp /tmp/orig.so dynamask_b200/lib/libdynamask_sm100.so
