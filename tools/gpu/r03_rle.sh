# parity tests, then the configs part of the bench (C3 / C4 / C5) and the paste timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r03d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r03d_pytest.log
python tools/gpu/paste_bench.py
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-competitor > gpurun_out/r03d_bench.json 2> gpurun_out/r03d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r03d_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r03d_bench.json'))
print(d['value'], d['ms_per_step'])
for k in ('c4','c5'):
    c=d['configs'][k]; print(k, c['img_per_s'], c['ms_per_image_this_rank'], c.get('ms_per_image_unpipelined'))
print('c3', d['configs']['c3']['bitmap_gt'], d['configs']['c3']['polygon_gt'])
print(d['extras']['dm_paste_masks']); print(d['extras']['inference_tail_per_image'])
PY
