# round 2, session 9: byte-lane strip path (witness), 16-bit Bayer tests, full GPU suite, default bench
mkdir -p gpurun_out
python tools/dbg_strip.py > gpurun_out/r2i_strip.log 2>&1; grep -c "OK" gpurun_out/r2i_strip.log; grep DIFF gpurun_out/r2i_strip.log | head -5
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2i_pytest.log 2>&1; tail -3 gpurun_out/r2i_pytest.log
timeout 400 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2i_bench.json'))
    print('bench', round(d['value']), d['ms_per_step'], d['config'].get('kernel_ms_per_step'), 'witness', d['roofline']['witness_debayer_gamma'], 'e2e', d['e2e'], d.get('parity'))
except Exception as e: print('ERR', e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_fused_strip' -s 3 -c 1 -f -o gpurun_out/r2i_strip4 python bench.py --frames 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2i_ncu4.log 2>&1; echo "ncu4 rc=$?"
ls -la gpurun_out | tail -6
