mkdir -p gpurun_out
timeout 900 python tools/stage_set_timing.py > gpurun_out/r2s_stage_sets.jsonl 2> gpurun_out/r2s_stage_sets.err; cat gpurun_out/r2s_stage_sets.jsonl; tail -3 gpurun_out/r2s_stage_sets.err
