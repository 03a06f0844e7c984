mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2af_bench_8gpu.json 2> gpurun_out/r2af_bench_8gpu.err; echo "bench 8 rc=$?"
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2af_bench_8gpu.json') if l.startswith('{')][-1]
print('N=8 value', round(d['value']), 'e2e', round(d['e2e']['value']), 'd2h_gbs', d['e2e'].get('d2h_gbs'), 'frac', d['e2e'].get('frac_of_copy_ceiling'), 'ms', round(d['ms_per_step'],3), 'parity', d['parity'] and d['parity']['differing_values'])
PY
