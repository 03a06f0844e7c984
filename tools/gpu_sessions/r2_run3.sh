mkdir -p gpurun_out
python tools/dbg_chain_dev.py > gpurun_out/r2d_chain_dev.log 2>&1; echo "chain_dev ok: $(grep -c 'differing pixels 0 ' gpurun_out/r2d_chain_dev.log) / 10"; grep "differing" gpurun_out/r2d_chain_dev.log | grep -v "pixels 0 " | head -3
python tools/dbg_strip.py > gpurun_out/r2d_strip.log 2>&1; echo "strip ok: $(grep -c OK gpurun_out/r2d_strip.log) / 40"; grep DIFF gpurun_out/r2d_strip.log | head -5; tail -3 gpurun_out/r2d_strip.log | grep -i error
python tools/dbg_strip.py 163 400 > gpurun_out/r2d_strip_b.log 2>&1; echo "strip (163 rows) ok: $(grep -c OK gpurun_out/r2d_strip_b.log) / 40"; grep DIFF gpurun_out/r2d_strip_b.log | head -5
for v in 1 2 4; do
  RIP_B200_FUSED_KERNEL=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2d_bench_v$v.json 2> gpurun_out/r2d_bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2d_bench_v$v.json'))
    print('variant $v', round(d['value']), round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['config']['kernel_ms_per_step'].items()}, 'witness', round(d['roofline']['witness_debayer_gamma']['avg_launch_ms'],4), round(d['roofline']['witness_debayer_gamma']['frac_of_peak'],3), 'parity', d['parity']['differing_values'])
except Exception as e: print('ERR $v', e)
PY
done
RIP_B200_FUSED_KERNEL=2 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_fused' -s 3 -c 1 -f -o gpurun_out/r2d_strip31 python bench.py --frames 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-witness --no-parity > gpurun_out/r2d_ncu31.log 2>&1; echo "ncu31 rc=$?"
RIP_B200_FUSED_KERNEL=4 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_fused' -s 3 -c 1 -f -o gpurun_out/r2d_strip31_v3 python bench.py --frames 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-witness --no-parity > gpurun_out/r2d_ncu31v3.log 2>&1; echo "ncu31v3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_fused' -s 7 -c 1 -f -o gpurun_out/r2d_strip4 python bench.py --frames 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2d_ncu4.log 2>&1; echo "ncu4 rc=$?"
