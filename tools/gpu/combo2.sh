bash tools/gpu/quick.sh c2
python tools/gpu/paste_bench.py
