# parity tests + the bench line + kernel-only durations of the current build
# usage: bash tools/gpu/r02_quick.sh <tag> ["<cfg>" ...]
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()}, d['roofline']['frac'])
print(d['e2e']); print(d.get('gpu_competitor'))
print({k:(v.get('img_per_s') if isinstance(v,dict) else None) for k,v in d.get('configs',{}).items()}, d['configs']['c3']['bitmap_gt']['ms_per_step'])
PY
if [ $# -eq 0 ]; then set -- "DM_NOP=0"; fi
bash tools/gpu/r02_kt.sh ${TAG} "$@"
