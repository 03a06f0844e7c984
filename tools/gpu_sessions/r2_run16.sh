# round 2, final-code session: default bench (all legs), reference arm, launch list, full-size ncu capture of the step's kernels
mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; head -c 300 gpurun_out/r2p_bench.json; echo
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p_bench_ref.json 2> gpurun_out/r2p_bench_ref.err; head -c 200 gpurun_out/r2p_bench_ref.json; echo
timeout 400 python bench.py --config 5 > gpurun_out/r2p_bench_c5.json 2> gpurun_out/r2p_bench_c5.err; head -c 200 gpurun_out/r2p_bench_c5.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2p_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2p_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_fused_fast|k_remap_tile|k_pca_stats_fast' -s 9 -c 3 -f \
  -o gpurun_out/r2p_step64 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-witness --no-parity > gpurun_out/r2p_step64.log 2>&1; echo "step capture rc=$?"
