#!/bin/bash
# quick GPU check: kernel parity tests + micro-benchmarks of the stage kernel (+ optional ncu capture: PROF=1)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_vp_system.py 2>&1 | tail -n 25 > gpurun_out/pytest_k.log
timeout 600 python -m pytest tests/test_gpu_vp_system.py -m gpu -q --tb=short 2>&1 | tail -n 150 > gpurun_out/pytest_s.log
rm -f gpurun_out/mb.log
for args in "" "--mode rhs" "--order 6" "--strict" "128 128 128 128"; do
  timeout 300 python tools/microbench_rhs.py $args >> gpurun_out/mb.log 2>&1
done
if [ -n "$PROF" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_march -s 1 -c 1 -o gpurun_out/prof_march -f \
  python tools/microbench_rhs.py 128 128 128 128 --reps 1 > gpurun_out/ncu_full.log 2>&1
fi
tail -n 5 gpurun_out/pytest_k.log; tail -n 5 gpurun_out/pytest_s.log; cat gpurun_out/mb.log
