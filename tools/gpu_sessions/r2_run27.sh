mkdir -p gpurun_out
for g in 2 4 8 16 32 64; do
RIP_B200_REMAP_FRAME_GROUP=$g timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-witness --steps 10 > gpurun_out/r2ad_bench_g$g.json 2> gpurun_out/r2ad_bench_g$g.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2ad_bench_g$g.json'))
print('group $g', round(d['value']), d['ms_per_step'], d['config']['kernel_ms_per_step']['remap'], d['parity']['max_abs_diff'])
PY
done
