mkdir -p gpurun_out
for c in 2 4; do timeout 400 python bench.py --config $c > gpurun_out/r2x_bench_c$c.json 2> gpurun_out/r2x_bench_c$c.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2x_bench_c$c.json'))
print('config $c', round(d['value']), d['e2e'].get('value'), d.get('latency_us'), d.get('parity',{}).get('max_abs_diff'))
PY
tail -2 gpurun_out/r2x_bench_c$c.err
done
