# run the GPU parity tests against library variants (dynamask_b200/lib/variants/*.so copied over the shipped name)
mkdir -p gpurun_out
cp dynamask_b200/lib/libdynamask_sm100.so /tmp/cur.so
for v in "$@"; do
  if [ "$v" = cur ]; then cp /tmp/cur.so dynamask_b200/lib/libdynamask_sm100.so; else cp dynamask_b200/lib/variants/$v.so dynamask_b200/lib/libdynamask_sm100.so; fi
  CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r03var_${v}.log 2>&1
  echo "== $v rc=$?"; grep -E "passed|failed|^FAILED|Error" gpurun_out/r03var_${v}.log | head -5
done
cp /tmp/cur.so dynamask_b200/lib/libdynamask_sm100.so
