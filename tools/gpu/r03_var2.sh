# parity tests + kernel-only durations per library variant
mkdir -p gpurun_out
cp dynamask_b200/lib/libdynamask_sm100.so /tmp/cur.so
for v in "$@"; do
  if [ "$v" = cur ]; then cp /tmp/cur.so dynamask_b200/lib/libdynamask_sm100.so; else cp dynamask_b200/lib/variants/$v.so dynamask_b200/lib/libdynamask_sm100.so; fi
  timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r03v2_${v}.log 2>&1
  rc=$?; echo "== $v rc=$rc"; grep -E "passed|failed|^FAILED|Error" gpurun_out/r03v2_${v}.log | head -5
  if [ $rc -eq 0 ]; then bash tools/gpu/r02_kt.sh r03v2_${v} "DM_NOP=0"; fi
done
cp /tmp/cur.so dynamask_b200/lib/libdynamask_sm100.so
