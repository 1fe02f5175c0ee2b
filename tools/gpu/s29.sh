python -m pytest tests -m gpu -x -q -k "roi_align or extractor or backward or full_size or simple or c3 or c5 or golden" > gpurun_out/s29_pytest.log 2>&1; tail -2 gpurun_out/s29_pytest.log
for d in 1 0 1 0; do
DM_RA_DYNAMIC=$d python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/s29.json 2> gpurun_out/s29.err
python -c "
import json; d=json.load(open('gpurun_out/s29.json')); print('dynamic $d', round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['kernels'].items()}, d['checksums'])" 2>&1 | tail -1
done
