#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sleep.log
for r in 1 2 3; do
for v in "" _sleep24 _sleep128 _sleep256; do
  LOKI_B200_LIB=loki_b200/libloki_b200$v.so timeout 300 python tools/microbench_rhs.py --reps 8 --fold >> gpurun_out/sleep.log 2>&1
done
done
for v in "" _sleep24 _sleep128 _sleep256; do
  LOKI_B200_LIB=loki_b200/libloki_b200$v.so timeout 300 python tools/microbench_rhs.py --reps 6 --fold --order 6 >> gpurun_out/sleep.log 2>&1
done
cat gpurun_out/sleep.log
