mkdir -p gpurun_out
timeout 400 python bench.py --no-cpu-baseline --no-e2e --no-witness > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2w_bench.json'))
print('bench', round(d['value']), d['ms_per_step'], d['config'].get('kernel_ms_per_step'), d['parity']['max_abs_diff'])
PY
timeout 400 python bench.py --config 5 --no-cpu-baseline --no-e2e > gpurun_out/r2w_bench_c5.json 2> gpurun_out/r2w_bench_c5.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2w_bench_c5.json'))
print('config5', round(d['value']), d['ms_per_step'], d['config'].get('kernel_ms_per_step'), d['parity']['max_abs_diff'])
PY
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2w_pytest.log 2>&1; tail -3 gpurun_out/r2w_pytest.log
