mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "ccc or CCC or golden or config5 or 4k or bench_size" > gpurun_out/r2r_pytest_ccc.log 2>&1; tail -3 gpurun_out/r2r_pytest_ccc.log
timeout 400 python bench.py --config 5 --no-cpu-baseline > gpurun_out/r2r_bench_c5.json 2> gpurun_out/r2r_bench_c5.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2r_bench_c5.json'))
print('config 5', round(d['value']), d['ms_per_step'], d['config'].get('kernel_ms_per_step'), d['e2e'].get('value'), d.get('ccc'), d['parity']['max_abs_diff'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2r_launches_c5.csv \
  python bench.py --config 5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2r_launches_c5.log 2>&1; echo "launch list rc=$?"
