mkdir -p gpurun_out
python tools/dbg_chain_dev.py > gpurun_out/r2b_chain_dev.log 2>&1; grep -c "differing pixels 0 " gpurun_out/r2b_chain_dev.log; grep "differing" gpurun_out/r2b_chain_dev.log | grep -v "pixels 0 " | head
python tools/dbg_strip.py > gpurun_out/r2b_strip.log 2>&1; grep -c "OK" gpurun_out/r2b_strip.log; grep DIFF gpurun_out/r2b_strip.log | head -5
for v in 0 1 2; do
  RIP_B200_FUSED_KERNEL=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_v$v.json 2> gpurun_out/r2b_bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2b_bench_v$v.json'))
    print('variant $v', round(d['value']), d['ms_per_step'], d['config']['kernel_ms_per_step'], 'witness', d['roofline']['witness_debayer_gamma']['avg_launch_ms'], d['roofline']['witness_debayer_gamma']['frac_of_peak'], d.get('parity'))
except Exception as e: print('ERR $v', e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_fused' -s 3 -c 1 -f -o gpurun_out/r2b_strip31 python bench.py --frames 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-witness --no-parity > gpurun_out/r2b_ncu31.log 2>&1; echo "ncu31 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_fused' -s 7 -c 1 -f -o gpurun_out/r2b_strip4 python bench.py --frames 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2b_ncu4.log 2>&1; echo "ncu4 rc=$?"
ls -la gpurun_out | tail -12
