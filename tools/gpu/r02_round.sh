# Round-2 evidence in one call: GPU tests, bench (both arms), launch list + per-kernel metrics of the
# bench command, kernel-only per-bucket / per-C3-call durations, --set full captures of the 14x14 and
# 28x28 single-bucket launches.
# usage: bash tools/gpu/r02_round.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
bash tools/gpu/round_profile.sh ${TAG} > gpurun_out/${TAG}_round.log 2>&1
bash tools/gpu/r02_kt.sh ${TAG}kt "DM_NOP=0" "DM_RA_BWD_DYNAMIC=1" > gpurun_out/${TAG}_kt.log 2>&1; cat gpurun_out/${TAG}_kt.log
bash tools/gpu/r02_smallprof.sh ${TAG} > gpurun_out/${TAG}_small.log 2>&1; tail -4 gpurun_out/${TAG}_small.log
head -c 1500 gpurun_out/${TAG}_bench.json; echo; cat gpurun_out/${TAG}_bench_ref.json | head -c 1500
