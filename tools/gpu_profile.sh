#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of the bench command + one full-set capture of the three hot kernels.
# usage: tools/gpu_profile.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-witness > gpurun_out/launches_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_fused|k_remap|k_pca_stats' -s 9 -c 3 -f \
  -o gpurun_out/prof_${TAG} python bench.py --frames 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-witness > gpurun_out/prof_${TAG}.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out
