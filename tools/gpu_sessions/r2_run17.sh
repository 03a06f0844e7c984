mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2q_launches_c5.csv \
  python bench.py --config 5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2q_launches_c5.log 2>&1; echo "launch list rc=$?"
