N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12
for m in threads procs; do timeout 60 ./tools/pcie_probe/probe --gpus $N --mode $m --seconds 2; done | tee gpurun_out/r2z_pcie_probe_${N}gpu_box.jsonl
timeout 60 ./tools/pcie_probe/probe --gpus 1 --mode threads --seconds 2 | tee -a gpurun_out/r2z_pcie_probe_${N}gpu_box.jsonl
