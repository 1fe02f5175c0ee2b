mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "paste or seg or smoke" 2>&1 | tail -2
for v in 1 4; do DM_PASTE_TILES=$v python tools/gpu/paste_bench.py; done 2>&1 | tee gpurun_out/paste_exp.log
