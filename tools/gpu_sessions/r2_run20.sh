mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()"
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2t_pytest.log 2>&1; tail -3 gpurun_out/r2t_pytest.log
