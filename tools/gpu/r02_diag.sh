TAG=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  name=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 300 python tools/bucket_breakdown.py > gpurun_out/${TAG}_buckets_${name}.json 2> gpurun_out/${TAG}_buckets_${name}.err
  echo "== $cfg"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_buckets_${name}.json'))
    print(' '.join('%s f %.3f b %.3f' % (k, v['fwd_ms'], v['bwd_ms_incl_zero_init']) for k, v in d.items()))
except Exception as e:
    print('failed', e)
PY
  tail -3 gpurun_out/${TAG}_buckets_${name}.err
done
