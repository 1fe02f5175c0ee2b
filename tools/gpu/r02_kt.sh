# Kernel-only durations (ncu gpu__time_duration, no host launch gaps) of the RoIAlign launches of
# tools/bucket_breakdown.py and tools/c3_breakdown.py for each env config.
# usage: bash tools/gpu/r02_kt.sh <tag> [cfg ...]
TAG=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  name=$(echo $cfg | tr ' =/' '___')
  for prog in bucket_breakdown c3_breakdown; do
    env $cfg timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ra_kernel --csv --log-file gpurun_out/${TAG}_${prog}_${name}.csv python tools/$prog.py > /dev/null 2> gpurun_out/${TAG}_${prog}_${name}.err
  done
  echo "== $cfg"; python tools/ncu_kt.py gpurun_out/${TAG}_bucket_breakdown_${name}.csv buckets; python tools/ncu_kt.py gpurun_out/${TAG}_c3_breakdown_${name}.csv c3
done
