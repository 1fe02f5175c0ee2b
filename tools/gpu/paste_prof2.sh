mkdir -p gpurun_out
python tools/gpu/paste_bench.py > gpurun_out/paste_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:paste_window -s 3 -c 1 -f -o gpurun_out/paste_win python tools/gpu/paste_bench.py > gpurun_out/paste_win_ncu.log 2>&1
tail -1 gpurun_out/paste_win_ncu.log
