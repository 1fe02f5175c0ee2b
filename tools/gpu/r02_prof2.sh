# --set full capture of the P14 / P28 forward launches of tools/bucket_breakdown.py (TMA path)
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ra_kernel -s 8 -c 1 -f -o gpurun_out/${TAG}_p14 python tools/bucket_breakdown.py > gpurun_out/${TAG}_ncu_p14.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ra_kernel -s 20 -c 1 -f -o gpurun_out/${TAG}_p28 python tools/bucket_breakdown.py > gpurun_out/${TAG}_ncu_p28.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_p14.log
