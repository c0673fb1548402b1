#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
for round in 1 2 3; do
  for lib in libloki_b200.so libloki_b200_s64lean.so libloki_b200_s200.so libloki_b200_s500.so libloki_b200_s20lean.so; do
    LOKI_B200_LIB=$PWD/loki_b200/$lib timeout 300 python tools/microbench_rhs.py --reps 8 >> gpurun_out/ab.log 2>&1
  done
done
cat gpurun_out/ab.log
