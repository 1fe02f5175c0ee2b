python -m pytest tests/test_configs_gpu.py -m gpu -x -q > gpurun_out/s13_pytest.log 2>&1; tail -3 gpurun_out/s13_pytest.log
python bench.py --no-cpu-baseline --no-extras > gpurun_out/s13_bench.json 2> gpurun_out/s13_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s13_bench.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
tail -3 gpurun_out/s13_bench.err
