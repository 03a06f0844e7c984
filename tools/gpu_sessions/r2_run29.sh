mkdir -p gpurun_out
python tools/dbg_strip.py > gpurun_out/r2ag_strip.log 2>&1; grep -c "OK" gpurun_out/r2ag_strip.log; grep DIFF gpurun_out/r2ag_strip.log | head -5
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2ag_pytest.log 2>&1; tail -3 gpurun_out/r2ag_pytest.log
timeout 600 python tools/stage_set_timing.py 2>/dev/null | grep -E '"wb|"gamma|"none' 
