set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
for v in 0 1 2; do
  RIP_B200_FUSED_KERNEL=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2a_bench_v$v.json 2> gpurun_out/r2a_bench_v$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2a_bench_v$v.json'))
print('variant $v', d['value'], d['ms_per_step'], d['config']['kernel_ms_per_step'], 'witness', d['roofline']['witness_debayer_gamma']['avg_launch_ms'], d['roofline']['witness_debayer_gamma']['frac_of_peak'], 'same', d['config']['device_equals_host_path'])
"
done
