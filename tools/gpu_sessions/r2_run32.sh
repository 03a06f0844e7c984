mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_apply.py -x -q -m gpu > gpurun_out/r2aj_pytest_apply.log 2>&1; tail -2 gpurun_out/r2aj_pytest_apply.log
for c in 2 4; do timeout 400 python bench.py --config $c --no-cpu-baseline > gpurun_out/r2aj_bench_c$c.json 2> gpurun_out/r2aj_bench_c$c.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2aj_bench_c$c.json'))
l=d.get('latency_us') or {}
print('config $c', round(d['value']), round(d['e2e']['value']), {k:(round(v['p50']),round(v['p99'])) for k,v in l.items() if isinstance(v,dict)}, d['parity']['max_abs_diff'])
PY
done
