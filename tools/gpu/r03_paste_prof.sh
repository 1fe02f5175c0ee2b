# --set full capture (with source) of the fused paste launch on the C4 shape
mkdir -p gpurun_out
python tools/gpu/paste_bench.py > gpurun_out/r03_paste_plain.log 2>&1; cat gpurun_out/r03_paste_plain.log
ncu --set full --clock-control none --import-source on -k regex:paste_fused -s 3 -c 1 -f -o gpurun_out/r03_paste python tools/gpu/paste_bench.py > gpurun_out/r03_paste_ncu.log 2>&1
tail -1 gpurun_out/r03_paste_ncu.log
