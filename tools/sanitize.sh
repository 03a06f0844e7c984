#!/bin/bash
# compute-sanitizer over every kernel family (run on a GPU box: `gpurun -- bash tools/sanitize.sh r2`); logs land in
# gpurun_out/ and are copied to profiles/<tag>_sanitize_*.log by hand after review.
TAG=${1:-r2}
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 3 --log-file gpurun_out/${TAG}_sanitize_${tool}.log \
    python tools/sanitize_workload.py > gpurun_out/${TAG}_sanitize_${tool}.out 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/${TAG}_sanitize_${tool}.out; tail -3 gpurun_out/${TAG}_sanitize_${tool}.log
done
