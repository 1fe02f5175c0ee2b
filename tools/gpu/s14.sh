# exact per-bucket kernel times: bucket_breakdown under ncu (gpu__time_duration only)
python tools/bucket_breakdown.py > gpurun_out/s14_plain.json 2> gpurun_out/s14_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ra_kernel --csv --log-file gpurun_out/s14_launches.csv python tools/bucket_breakdown.py > gpurun_out/s14_ncu.log 2>&1
cat gpurun_out/s14_plain.json | head -60
