# 8-GPU session (run with gpurun --gpus 8): copy ceiling at 1/2/4/8 GPUs, scaling of the bench at 2/4/8, configs 4 and 5,
# the in-process multi-GPU tests.  Everything bounded by timeouts.
mkdir -p gpurun_out
nvidia-smi -L | head -8
for g in 1 2 4 8; do for m in threads procs; do timeout 60 ./tools/pcie_probe/probe --gpus $g --mode $m --seconds 2; done; done | tee gpurun_out/r2_pcie_probe.jsonl
timeout 60 ./tools/pcie_probe/probe --gpus 8 --mode procs --seconds 2 --dir d2h | tee -a gpurun_out/r2_pcie_probe.jsonl
timeout 60 ./tools/pcie_probe/probe --gpus 8 --mode procs --seconds 2 --dir h2d | tee -a gpurun_out/r2_pcie_probe.jsonl
timeout 600 python -m pytest tests -m gpu -q -k "two_gpus or multi_gpu" > gpurun_out/r2_pytest_multi_gpu.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r2_pytest_multi_gpu.log
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench_${n}gpu.json 2> gpurun_out/r2_bench_${n}gpu.err; echo "bench $n rc=$?"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r2_bench_${n}gpu.json') if l.startswith('{')][-1]
    print('N=$n value', round(d['value']), 'e2e', round(d['e2e']['value']), 'd2h_gbs', d['e2e'].get('d2h_gbs'), 'ms', round(d['ms_per_step'],3), 'parity', d['parity'] and d['parity']['differing_values'])
except Exception as e: print('ERR', e)
PY
done
for c in 4 5; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --config $c --gpus 8 --steps 5 --warmup 3 --e2e-steps 3 > gpurun_out/r2_bench_c${c}_8gpu.json 2> gpurun_out/r2_bench_c${c}_8gpu.err; echo "config $c x8 rc=$?"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r2_bench_c${c}_8gpu.json') if l.startswith('{')][-1]
    print('config $c N=8 value', round(d['value']), 'e2e', round(d['e2e']['value']), 'parity', d['parity'] and d['parity']['differing_values'], [ (s['rank'], s['latency_us'] and round(s['latency_us']['p50']), s['latency_us'] and round(s['latency_us']['p99'])) for s in d.get('streams', [])], d.get('ccc'))
except Exception as e: print('ERR', e)
PY
done
# one process, 8 GPUs: the product entry point (rip_apply_batch_host_multi)
timeout 300 python tools/multi_gpu_bench.py > gpurun_out/r2_multi_gpu_entry.json 2> gpurun_out/r2_multi_gpu_entry.err; tail -2 gpurun_out/r2_multi_gpu_entry.json
