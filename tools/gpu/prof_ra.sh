# --set full capture (with source) of one forward and one backward RoIAlign launch.  usage: prof_ra.sh <tag>
TAG=${1:-ra}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ra_kernel -s 2 -c 2 -f -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
