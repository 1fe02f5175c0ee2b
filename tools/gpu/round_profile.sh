# Round-end evidence: full bench line, reference arm, ncu launch list and per-kernel DRAM metrics.
# usage: bash tools/gpu/round_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-competitor"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > /dev/null 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem --clock-control none -k regex:'ra_kernel|paste|assign|mask_target|polygon|refine|rle' --csv --log-file gpurun_out/${TAG}_kernel_metrics.csv $CMD > /dev/null 2>&1
cat gpurun_out/${TAG}_bench.json | head -c 3000; echo; cat gpurun_out/${TAG}_bench_ref.json
