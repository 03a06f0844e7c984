# last session of the round: the bench lines of the final commit (configs 3, 2, 4, 5 at one GPU)
mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/r2ai_bench.json 2> gpurun_out/r2ai_bench.err; head -c 200 gpurun_out/r2ai_bench.json; echo
for c in 2 4 5; do timeout 400 python bench.py --config $c > gpurun_out/r2ai_bench_c$c.json 2> gpurun_out/r2ai_bench_c$c.err; head -c 160 gpurun_out/r2ai_bench_c$c.json; echo; done
