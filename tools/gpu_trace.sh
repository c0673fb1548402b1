#!/bin/bash
mkdir -p gpurun_out
LOKI_B200_LIB=$PWD/loki_b200/libloki_b200_trace.so timeout 300 python tools/microbench_rhs.py --reps 1 > gpurun_out/trace.log 2>&1
grep -c "^TR" gpurun_out/trace.log; tail -2 gpurun_out/trace.log
