# usage: bash tools/gpu/r02_tma2.sh <tag> [cfg ...]   -- parity, then bucket / C3 timings for each env config
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for cfg in "$@"; do
  name=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 300 python tools/bucket_breakdown.py > gpurun_out/${TAG}_buckets_${name}.json 2> gpurun_out/${TAG}_buckets_${name}.err
  env $cfg timeout 300 python tools/c3_breakdown.py > gpurun_out/${TAG}_c3_${name}.json 2>> gpurun_out/${TAG}_buckets_${name}.err
  echo "== $cfg"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_buckets_${name}.json'))
    print(' '.join('%s f %.3f b %.3f' % (k, v['fwd_ms'], v['bwd_ms_incl_zero_init']) for k, v in d.items()))
    d=json.load(open('gpurun_out/${TAG}_c3_${name}.json'))
    print(' '.join('%s f %.3f b %.3f' % (k, v['fwd_ms'], v['bwd_ms_incl_zero_init']) for k, v in d.items() if isinstance(v, dict)))
except Exception as e:
    print('failed', e)
PY
  tail -3 gpurun_out/${TAG}_buckets_${name}.err
done
