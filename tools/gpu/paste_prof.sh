mkdir -p gpurun_out
python tools/gpu/paste_bench.py > gpurun_out/paste_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:paste_kernel -s 3 -c 1 -f -o gpurun_out/paste_v4 python tools/gpu/paste_bench.py > gpurun_out/paste_v4_ncu.log 2>&1
cat gpurun_out/paste_prof_plain.log
