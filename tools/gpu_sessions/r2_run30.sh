mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_apply.py -x -q -m gpu > gpurun_out/r2ah_pytest_apply.log 2>&1; tail -3 gpurun_out/r2ah_pytest_apply.log
timeout 400 python bench.py --config 2 --no-cpu-baseline > gpurun_out/r2ah_bench_c2.json 2> gpurun_out/r2ah_bench_c2.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2ah_bench_c2.json'))
print('config 2', d.get('latency_us'), d.get('parity',{}).get('max_abs_diff'))
PY
tail -2 gpurun_out/r2ah_bench_c2.err
