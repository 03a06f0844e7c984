# second 8-GPU session (gpurun --gpus 8), kept short: the product multi-GPU entry with on-demand chunk dealing, the in-process tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "two_gpus or multi_gpu" > gpurun_out/r2_pytest_multi_gpu.log 2>&1; echo "pytest multi rc=$?"; tail -2 gpurun_out/r2_pytest_multi_gpu.log
ONLY_GPUS=8 FRAMES_PER_GPU=64 timeout 300 python tools/multi_gpu_bench.py > gpurun_out/r2_multi_gpu_entry.json 2> gpurun_out/r2_multi_gpu_entry.err; tail -2 gpurun_out/r2_multi_gpu_entry.json; tail -3 gpurun_out/r2_multi_gpu_entry.err
