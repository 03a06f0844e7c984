mkdir -p gpurun_out
python - <<PY
import sys, numpy as np
sys.path.insert(0,'tests')
import conftest
from test_gpu_parity import make_pair, FULL
rows, cols = 480, 640
rng = np.random.default_rng(5)
frames = rng.integers(0, 256, (6, rows, cols), dtype=np.uint8)
nbad = 0
for rep in range(6):
  for first_pin in (False, True):
    for pinned in (False, True):
      p,o = make_pair(rows, cols, **FULL)
      if first_pin: pin = p.pinned_empty((rows, cols))
      p.use_pinned_results = pinned
      ref = o.apply(frames[rep], "bayer_bggr8")[0]
      for it in range(2):
        got = p.process(frames[rep], "bayer_bggr8")
        bad = np.argwhere((got != ref).any(axis=2))
        nbad += len(bad)
        if len(bad): print('rep', rep, 'first_pin', first_pin, 'pinned_out', pinned, 'iter', it, 'bad px', len(bad), 'rows', (bad[:,0].min(), bad[:,0].max()))
print('first-frame check: total bad px', nbad)
PY
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2m_pytest.log 2>&1; tail -3 gpurun_out/r2m_pytest.log
for c in 2 4; do timeout 400 python bench.py --config $c > gpurun_out/r2m_bench_c$c.json 2> gpurun_out/r2m_bench_c$c.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2m_bench_c$c.json'))
print('config $c', round(d['value']), d['e2e'].get('value'), d.get('latency_us'), d.get('parity',{}).get('max_abs_diff'))
PY
done
python tools/apply_latency.py > gpurun_out/r2m_apply_latency.log 2>&1; tail -4 gpurun_out/r2m_apply_latency.log
