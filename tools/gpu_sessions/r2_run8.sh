mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest.log
for c in 2 4 5; do
  timeout 400 python bench.py --config $c --steps 5 --warmup 3 --e2e-steps 3 > gpurun_out/r2h_bench_c$c.json 2> gpurun_out/r2h_bench_c$c.err; echo "config $c rc=$?"; tail -3 gpurun_out/r2h_bench_c$c.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2h_bench_c$c.json'))
    print('config $c value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), 'parity', d['parity'] and d['parity']['differing_values'], d.get('latency_us'), d.get('ccc'))
except Exception as e: print('ERR $c', e)
PY
done
