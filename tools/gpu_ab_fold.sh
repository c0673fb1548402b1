#!/bin/bash
# round 2, call E: velocity-boundary fold (rolled, out-of-line) -- parity, A/B against the kernel without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipe.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_pipe.log
rm -f gpurun_out/ab.log
for round in 1 2; do
  timeout 300 python tools/microbench_rhs.py --reps 8 >> gpurun_out/ab.log 2>&1
  timeout 300 python tools/microbench_rhs.py --reps 8 --fold >> gpurun_out/ab.log 2>&1
  LOKI_B200_LIB=loki_b200/libloki_b200_nofold.so timeout 300 python tools/microbench_rhs.py --reps 8 >> gpurun_out/ab.log 2>&1
  LOKI_B200_LIB=loki_b200/libloki_b200_nofold.so timeout 300 python tools/microbench_rhs.py --reps 8 --fold >> gpurun_out/ab.log 2>&1
  timeout 300 python tools/microbench_rhs.py --reps 8 --order 6 >> gpurun_out/ab.log 2>&1
  timeout 300 python tools/microbench_rhs.py --reps 8 --order 6 --fold >> gpurun_out/ab.log 2>&1
  LOKI_B200_LIB=loki_b200/libloki_b200_nofold.so timeout 300 python tools/microbench_rhs.py --reps 8 --order 6 >> gpurun_out/ab.log 2>&1
done
cat gpurun_out/pytest_pipe.log gpurun_out/ab.log
