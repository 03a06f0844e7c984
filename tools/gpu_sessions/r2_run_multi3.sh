# third 8-GPU session: configs[3] (8 camera streams, one per GPU, per-frame process()) with the copy pool sized by ranks per
# host (default) against the old fixed 8 copy threads per rank; then the default bench at 8 GPUs with the final code
mkdir -p gpurun_out
for mode in default fixed8; do
  if [ $mode = fixed8 ]; then export RIP_B200_COPY_THREADS=8; else unset RIP_B200_COPY_THREADS; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --config 4 --gpus 8 --steps 5 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/r2u_bench_c4_8gpu_$mode.json 2> gpurun_out/r2u_bench_c4_8gpu_$mode.err; echo "config 4 x8 $mode rc=$?"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r2u_bench_c4_8gpu_$mode.json') if l.startswith('{')][-1]
    print('$mode value', round(d['value']), 'e2e', round(d['e2e']['value']), 'parity', d['parity'] and d['parity']['differing_values'], [ (s['rank'], s['latency_us'] and round(s['latency_us']['p50']), s['latency_us'] and round(s['latency_us']['p99'])) for s in d.get('streams', [])])
except Exception as e: print('ERR', e)
PY
done
unset RIP_B200_COPY_THREADS
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_8gpu.json 2> gpurun_out/r2u_bench_8gpu.err; echo "bench 8 rc=$?"
python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r2u_bench_8gpu.json') if l.startswith('{')][-1]
    print('N=8 value', round(d['value']), 'e2e', round(d['e2e']['value']), 'd2h_gbs', d['e2e'].get('d2h_gbs'), 'frac', d['e2e'].get('frac_of_copy_ceiling'), 'ms', round(d['ms_per_step'],3), 'parity', d['parity'] and d['parity']['differing_values'])
except Exception as e: print('ERR', e)
PY
