# quick GPU check: parity tests (optionally filtered) + a short device-resident bench
# usage: bash tools/gpu/quick.sh <tag> [pytest -k expr]
TAG=${1:-q}
K=${2:-}
mkdir -p gpurun_out
if [ -n "$K" ]; then python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/${TAG}_pytest.log 2>&1; else python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; fi
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'])
for k,v in d['kernels'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in('ms','achieved_gbs','frac_of_measured_peak')})
for k,v in d.get('extras',{}).items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a!='workload'})
PY
tail -3 gpurun_out/${TAG}_bench.err
