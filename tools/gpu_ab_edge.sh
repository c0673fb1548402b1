#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/edge.log
timeout 300 python -m pytest tests/test_gpu_pipe.py -x -q -k "folds or equals_generic" 2>&1 | tail -2 >> gpurun_out/edge.log
LOKI_B200_LIB=loki_b200/libloki_b200_edgefirst.so timeout 300 python -m pytest tests/test_gpu_pipe.py -x -q -k "folds or equals_generic" 2>&1 | tail -2 >> gpurun_out/edge.log
for r in 1 2 3; do
for v in "" _edgefirst; do
  LOKI_B200_LIB=loki_b200/libloki_b200$v.so timeout 300 python tools/microbench_rhs.py --reps 8 --fold >> gpurun_out/edge.log 2>&1
done
done
cat gpurun_out/edge.log
