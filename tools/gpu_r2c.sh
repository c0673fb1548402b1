#!/bin/bash
# round 2, call C: pipe v3 parity + A/B + bench line + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipe.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_pipe.log
rm -f gpurun_out/ab.log
for round in 1 2; do
  timeout 300 python tools/microbench_rhs.py --reps 8 >> gpurun_out/ab.log 2>&1
  timeout 300 python tools/microbench_rhs.py --reps 8 --variant 2 >> gpurun_out/ab.log 2>&1
  timeout 300 python tools/microbench_rhs.py --reps 8 --order 6 >> gpurun_out/ab.log 2>&1
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_pipe -s 1 -c 1 -o gpurun_out/prof_pipe -f \
  python tools/microbench_rhs.py 256 256 128 128 --reps 1 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/pytest_pipe.log gpurun_out/ab.log; tail -2 gpurun_out/bench.log; tail -3 gpurun_out/bench.err; tail -3 gpurun_out/ncu_full.log
