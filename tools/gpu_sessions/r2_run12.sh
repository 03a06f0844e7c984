mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_apply.py -x -q -m gpu -k pinned > gpurun_out/r2l_pytest_pinned.log 2>&1; tail -30 gpurun_out/r2l_pytest_pinned.log
python - <<PY
import sys, numpy as np
sys.path.insert(0,'tests')
import conftest
from test_gpu_parity import make_pair, FULL
rows, cols = 480, 640
rng = np.random.default_rng(5)
frames = rng.integers(0, 256, (6, rows, cols), dtype=np.uint8)
for first_pin in (False, True):
  for pinned in (False, True):
    p,o = make_pair(rows, cols, **FULL)
    if first_pin: pin = p.pinned_empty((rows, cols))
    p.use_pinned_results = pinned
    ref = o.apply(frames[0], "bayer_bggr8")[0]
    for it in range(4):
        got = p.process(frames[0], "bayer_bggr8")
        bad = np.argwhere((got != ref).any(axis=2))
        print('first_pin', first_pin, 'pinned_out', pinned, 'iter', it, 'bad px', len(bad), 'rows', (bad[:,0].min(), bad[:,0].max()) if len(bad) else None, 'cols', (bad[:,1].min(), bad[:,1].max()) if len(bad) else None)
PY
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; tail -3 gpurun_out/r2l_pytest.log
timeout 400 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2l_bench.json'))
print('bench', round(d['value']), d['ms_per_step'], d['config'].get('kernel_ms_per_step'), 'witness', d['roofline']['witness_debayer_gamma']['avg_launch_ms'], d['parity']['max_abs_diff'])
PY
