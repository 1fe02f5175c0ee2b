mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "paste or seg or smoke" 2>&1 | tail -2
python tools/gpu/paste_bench.py > gpurun_out/paste_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -k regex:paste -s 6 -c 6 --csv --log-file gpurun_out/paste_launches.csv python tools/gpu/paste_bench.py > /dev/null 2>&1
cat gpurun_out/paste_prof_plain.log
python - <<'PY'
import csv,io
lines=open('gpurun_out/paste_launches.csv').read().splitlines()
i=[k for k,l in enumerate(lines) if l.startswith('"ID"')][0]
for r in csv.DictReader(io.StringIO('\n'.join(lines[i:]))):
    print(r['ID'], r['Kernel Name'][:50], r['Grid Size'], r['Metric Name'], r['Metric Value'])
PY
