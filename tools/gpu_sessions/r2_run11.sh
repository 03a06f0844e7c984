# round 2, session 11: pinned caller buffers in rip_apply, remap kernel with warp-uniform index, full suite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_pytest.log 2>&1; tail -3 gpurun_out/r2k_pytest.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2k_bench.json'))
print('bench', round(d['value']), d['ms_per_step'], d['config'].get('kernel_ms_per_step'), 'witness', d['roofline']['witness_debayer_gamma']['avg_launch_ms'], d['parity']['max_abs_diff'])
PY
for c in 2 4; do timeout 400 python bench.py --config $c --no-cpu-baseline > gpurun_out/r2k_bench_c$c.json 2> gpurun_out/r2k_bench_c$c.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2k_bench_c$c.json'))
print('config $c', round(d['value']), d['e2e'].get('value'), d.get('latency_us'), d.get('parity',{}).get('max_abs_diff'))
PY
done
python tools/apply_latency.py > gpurun_out/r2k_apply_latency.log 2>&1; tail -4 gpurun_out/r2k_apply_latency.log
