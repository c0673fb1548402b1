#!/bin/bash
# round 2, call A: fp64 ceiling, parity of the pipelined stage kernel, A/B of kernel variants on ONE box
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_r2a.csv &
SMI=$!
timeout 120 tools/fp64_peak 3 > gpurun_out/fp64_peak.log 2>&1
timeout 900 python -m pytest tests/test_gpu_pipe.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_pipe.log
rm -f gpurun_out/ab.log
for round in 1 2; do
  LOKI_B200_LIB=$PWD/loki_b200/libloki_b200.so timeout 300 python tools/microbench_rhs.py --reps 8 >> gpurun_out/ab.log 2>&1
  LOKI_B200_LIB=$PWD/loki_b200/libloki_b200.so timeout 300 python tools/microbench_rhs.py --reps 8 --variant 2 >> gpurun_out/ab.log 2>&1
  for gg in "8 4" "4 8" "6 6" "16 2" "32 1" "2 16"; do
    set -- $gg
    echo "pipe gy=$1 gv=$2" >> gpurun_out/ab.log
    LK_PIPE_GY=$1 LK_PIPE_GV=$2 timeout 300 python tools/microbench_rhs.py --reps 8 >> gpurun_out/ab.log 2>&1
  done
  for lib in g84 g48 g66; do
    LOKI_B200_LIB=$PWD/loki_b200/libloki_b200_$lib.so timeout 300 python tools/microbench_rhs.py --reps 8 >> gpurun_out/ab.log 2>&1
  done
  LOKI_B200_LIB=$PWD/loki_b200/libloki_b200.so timeout 300 python tools/microbench_rhs.py --reps 8 --order 6 >> gpurun_out/ab.log 2>&1
  LOKI_B200_LIB=$PWD/loki_b200/libloki_b200.so timeout 300 python tools/microbench_rhs.py --reps 8 --order 6 --variant 2 >> gpurun_out/ab.log 2>&1
done
kill $SMI
cat gpurun_out/fp64_peak.log gpurun_out/pytest_pipe.log gpurun_out/ab.log
