# round 2, session 10: full GPU suite, default bench, launch list, full-size ncu captures (traffic), other configs at N=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2j_pytest.log 2>&1; tail -3 gpurun_out/r2j_pytest.log
timeout 400 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; head -c 600 gpurun_out/r2j_bench.json; echo
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err; head -c 400 gpurun_out/r2j_bench_ref.json; echo
for c in 2 4 5; do timeout 400 python bench.py --config $c > gpurun_out/r2j_bench_c$c.json 2> gpurun_out/r2j_bench_c$c.err; head -c 300 gpurun_out/r2j_bench_c$c.json; echo; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2j_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2j_launches.log 2>&1; echo "launch list rc=$?"
# full-size (64 frames) captures: one launch each of the stats, fused and remap kernels of a timed step, then the witness
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_fused_fast|k_remap_tile|k_pca_stats_fast' -s 9 -c 3 -f \
  -o gpurun_out/r2j_step64 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-witness --no-parity > gpurun_out/r2j_step64.log 2>&1; echo "step capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_fused_strip' -s 2 -c 1 -f \
  -o gpurun_out/r2j_witness64 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2j_witness64.log 2>&1; echo "witness capture rc=$?"
python tools/apply_latency.py > gpurun_out/r2j_apply_latency.log 2>&1; tail -5 gpurun_out/r2j_apply_latency.log
ls -la gpurun_out | tail -14
