# One optimisation step on the GPU: parity tests, kernel-only durations (ncu) of the per-bucket and
# per-C3-call launches for each env config, optional --set full captures of C3 launches.
# usage: bash tools/gpu/r02_step.sh <tag> "<cfg>" ["<cfg>" ...] [-- <ra_kernel launch index of tools/c3_breakdown.py> ...]
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
CFGS=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do CFGS+=("$1"); shift; done
[ "$1" == "--" ] && shift
bash tools/gpu/r02_kt.sh ${TAG} "${CFGS[@]}"
for idx in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ra_kernel -s $idx -c 1 -f -o gpurun_out/${TAG}_c3_$idx python tools/c3_breakdown.py > gpurun_out/${TAG}_ncu_c3_$idx.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_c3_$idx.log
done
