#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_pipe -s 1 -c 1 -o gpurun_out/prof_pipe_o6 -f \
  python tools/microbench_rhs.py 256 256 128 128 --reps 1 --fold --order 6 > gpurun_out/ncu_full_o6.log 2>&1
tail -2 gpurun_out/ncu_full_o6.log
