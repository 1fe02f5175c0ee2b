# One --set full capture (with source) of the RoIAlign fwd/bwd kernels and of the paste kernel.
# usage: bash tools/gpu/profile_full.sh <tag>
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ra_kernel -s 2 -c 2 -f -o gpurun_out/${TAG}_ra $CMD > gpurun_out/${TAG}_ncu_ra.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:paste_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_paste $CMD > gpurun_out/${TAG}_ncu_paste.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mask_target_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_mt $CMD > gpurun_out/${TAG}_ncu_mt.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_ra.log
ls -la gpurun_out
