# Final round-2 evidence in one call: smoke(), GPU tests, bench (both arms), launch list + per-kernel metrics of
# the bench command, kernel-only per-bucket / per-C3-call durations, paste timing, per-image tail timing.
TAG=${1:-r04}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
bash tools/gpu/round_profile.sh ${TAG} > gpurun_out/${TAG}_round.log 2>&1
bash tools/gpu/r02_kt.sh ${TAG}kt "DM_NOP=0" > gpurun_out/${TAG}_kt.log 2>&1; cat gpurun_out/${TAG}_kt.log
python tools/gpu/paste_bench.py
python tools/gpu/r03_tail.py 32 | cut -c1-100
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_tail_launches.csv python tools/gpu/r03_tail.py 4 > /dev/null 2>&1
python profiles/summarize_ncu.py launches gpurun_out/${TAG}_tail_launches.csv | grep "dm::"
head -c 1200 gpurun_out/${TAG}_bench.json; echo; head -c 600 gpurun_out/${TAG}_bench_ref.json; echo
