#!/bin/bash
# one ncu --set full capture of the stage kernel at 128^4 (+ the micro-benchmark numbers of the same build)
mkdir -p gpurun_out; rm -f gpurun_out/mb.log
for args in "" "--mode rhs" "$@"; do
  timeout 300 python tools/microbench_rhs.py $args >> gpurun_out/mb.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_march -s 1 -c 1 -o gpurun_out/prof_march -f \
  python tools/microbench_rhs.py 128 128 128 128 --reps 1 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/mb.log; tail -n 2 gpurun_out/ncu_full.log
