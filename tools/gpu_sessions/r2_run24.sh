mkdir -p gpurun_out
python tools/dbg_strip.py > gpurun_out/r2aa_strip.log 2>&1; grep -c "OK" gpurun_out/r2aa_strip.log; grep DIFF gpurun_out/r2aa_strip.log | head -5
timeout 400 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2aa_bench.json'))
print('bench', round(d['value']), d['ms_per_step'], 'witness', d['roofline']['witness_debayer_gamma']['avg_launch_ms'], d['roofline']['witness_debayer_gamma']['frac_of_peak'], d['parity']['max_abs_diff'])
PY
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2aa_pytest.log 2>&1; tail -3 gpurun_out/r2aa_pytest.log
