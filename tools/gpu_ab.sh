#!/bin/bash
# A/B micro-benchmarks of kernel variants on ONE box: alternate the libraries, 2 rounds, 8 reps each
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
LIBS="loki_b200/libloki_b200.so $(ls loki_b200/libloki_b200_*.so 2>/dev/null)"
for round in 1 2; do
  for lib in $LIBS; do
    for args in "" "--order 6"; do
      LOKI_B200_LIB=$PWD/$lib timeout 300 python tools/microbench_rhs.py $args --reps 8 >> gpurun_out/ab.log 2>&1
    done
  done
done
cat gpurun_out/ab.log
