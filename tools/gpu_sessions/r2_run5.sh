mkdir -p gpurun_out
python tools/dbg_chain_dev.py > gpurun_out/r2f_chain_dev.log 2>&1; echo "chain_dev ok: $(grep -c 'differing pixels 0 ' gpurun_out/r2f_chain_dev.log) / 10"
python tools/dbg_strip.py > gpurun_out/r2f_strip.log 2>&1; echo "strip ok: $(grep -c OK gpurun_out/r2f_strip.log) / 40"; grep DIFF gpurun_out/r2f_strip.log | head -3
for v in 1 2 3; do
  RIP_B200_FUSED_KERNEL=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_v$v.json 2> gpurun_out/r2f_bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2f_bench_v$v.json'))
    print('variant $v', round(d['value']), round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['config']['kernel_ms_per_step'].items()}, 'witness', round(d['roofline']['witness_debayer_gamma']['avg_launch_ms'],4), round(d['roofline']['witness_debayer_gamma']['frac_of_peak'],3), 'parity', d['parity']['differing_values'])
except Exception as e: print('ERR $v', e)
PY
done
RIP_B200_FUSED_KERNEL=3 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_fused' -s 3 -c 1 -f -o gpurun_out/r2f_strip31_v3 python bench.py --frames 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-witness --no-parity > gpurun_out/r2f_ncu31v3.log 2>&1; echo "ncu31v3 rc=$?"
