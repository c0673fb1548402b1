#!/bin/bash
# quick GPU check: kernel parity tests + micro-benchmarks of the stage kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_vp_system.py 2>&1 | tail -25 > gpurun_out/pytest_k.log
timeout 600 python -m pytest tests/test_gpu_vp_system.py -m gpu -q --tb=short 2>&1 | tail -150 > gpurun_out/pytest_s.log
rm -f gpurun_out/mb.log
for args in "" "--mode rhs" "--order 6" "--variant 1" "--strict" "128 128 128 128"; do
  timeout 300 python tools/microbench_rhs.py $args >> gpurun_out/mb.log 2>&1
done
tail -5 gpurun_out/pytest_k.log; tail -5 gpurun_out/pytest_s.log; cat gpurun_out/mb.log
