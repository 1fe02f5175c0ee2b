# paste timing per library variant (dynamask_b200/lib/variants/*.so copied over the shipped name)
cp dynamask_b200/lib/libdynamask_sm100.so /tmp/cur.so
for v in "$@"; do
  if [ "$v" = cur ]; then cp /tmp/cur.so dynamask_b200/lib/libdynamask_sm100.so; else cp dynamask_b200/lib/variants/$v.so dynamask_b200/lib/libdynamask_sm100.so; fi
  echo "== $v"; python tools/gpu/paste_bench.py; DM_PASTE_DIAG=2 python tools/gpu/paste_bench.py
done
cp /tmp/cur.so dynamask_b200/lib/libdynamask_sm100.so
