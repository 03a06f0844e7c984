mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2ac_pytest.log 2>&1; tail -3 gpurun_out/r2ac_pytest.log
timeout 500 python bench.py > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err; head -c 250 gpurun_out/r2ac_bench.json; echo
timeout 400 python bench.py --config 5 > gpurun_out/r2ac_bench_c5.json 2> gpurun_out/r2ac_bench_c5.err; head -c 250 gpurun_out/r2ac_bench_c5.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2ac_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2ac_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_remap_tile' -s 3 -c 1 -f \
  -o gpurun_out/r2ac_remap64 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-witness --no-parity > gpurun_out/r2ac_remap64.log 2>&1; echo "remap capture rc=$?"
python tools/sanitize_workload.py | tail -1
