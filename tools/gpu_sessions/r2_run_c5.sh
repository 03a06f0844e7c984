# configs[4] (bench.py --config 5) under torchrun at the GPU count of the box
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --config 5 --gpus $N --steps 10 --warmup 3 --e2e-steps 5 --no-cpu-baseline > gpurun_out/r2y_bench_c5_${N}gpu.json 2> gpurun_out/r2y_bench_c5_${N}gpu.err; echo "config 5 x$N rc=$?"
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2y_bench_c5_${N}gpu.json') if l.startswith('{')][-1]
print('config 5 N=$N value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'frac', d['e2e'].get('frac_of_copy_ceiling'), 'parity', d['parity'] and d['parity']['differing_values'], d.get('ccc',{}).get('uv_last_frame'))
PY
