mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 1500 gpurun_out/r2g_bench.json
for g in 1; do for m in threads procs; do ./tools/pcie_probe/probe --gpus $g --mode $m --seconds 2; done; done
./tools/pcie_probe/probe --gpus 1 --mode threads --seconds 2 --dir d2h; ./tools/pcie_probe/probe --gpus 1 --mode threads --seconds 2 --dir h2d
