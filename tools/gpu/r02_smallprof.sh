# Round 2, step 0: where does the time go on the small pooled sizes?  Per-bucket and per-C3-call
# timings of the current code, then --set full captures (with source) of the 14x14 and 28x28
# single-bucket forward + backward launches of tools/bucket_breakdown.py.
# usage: bash tools/gpu/r02_smallprof.sh <tag>
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.log 2>&1
timeout 300 python tools/bucket_breakdown.py > gpurun_out/${TAG}_buckets.json 2> gpurun_out/${TAG}_buckets.err
timeout 300 python tools/c3_breakdown.py > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
cat gpurun_out/${TAG}_buckets.json gpurun_out/${TAG}_c3.json
# P14: ra_kernel launches 0..11 (fwd, bwd alternating); P28: 12..23
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ra_kernel -s 8 -c 2 -f -o gpurun_out/${TAG}_p14 python tools/bucket_breakdown.py > gpurun_out/${TAG}_ncu_p14.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ra_kernel -s 20 -c 2 -f -o gpurun_out/${TAG}_p28 python tools/bucket_breakdown.py > gpurun_out/${TAG}_ncu_p28.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_p14.log gpurun_out/${TAG}_ncu_p28.log
ls -la gpurun_out | tail -8
