# TMA forward: parity first, then A/B timings.  usage: bash tools/gpu/r02_tma1.sh <tag>
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
for cfg in "DM_RA_TMA=1" "DM_RA_TMA=0" "DM_RA_TMA=1 DM_RA_TMA_ROWMAJOR=0" "DM_RA_TMA=1 DM_RA_TMA_SLOTS=4" "DM_RA_TMA=1 DM_RA_TMA_SLOTS=2" "DM_RA_TMA=1 DM_RA_TMA_L2=2" "DM_RA_TMA=1 DM_RA_FWD_THREADS=128 DM_RA_FWD_SMEM_KB=52"; do
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
