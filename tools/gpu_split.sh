#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/split.log
for r in 1 2; do
for sp in 0 1 2 3; do timeout 300 python tools/microbench_rhs.py --reps 8 --fold --split $sp >> gpurun_out/split.log 2>&1; done
done
timeout 300 python tools/microbench_rhs.py --reps 8 --fold --split 3 --order 6 >> gpurun_out/split.log 2>&1
timeout 300 python tools/microbench_rhs.py --reps 8 --fold --order 6 >> gpurun_out/split.log 2>&1
cat gpurun_out/split.log
