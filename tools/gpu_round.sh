#!/bin/bash
# one gpurun call: parity tests, micro-benchmarks, bench line, ncu launch list and one full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest.log
rm -f gpurun_out/mb.log
for args in "" "--mode rhs" "--order 6" "--order 6 --mode rhs" "--variant 1" "--strict"; do
  timeout 300 python tools/microbench_rhs.py $args >> gpurun_out/mb.log 2>&1
done
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_stage_march --csv --log-file gpurun_out/traffic.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_march -s 1 -c 1 -o gpurun_out/prof_march -f \
  python tools/microbench_rhs.py 256 256 128 128 --reps 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest.log; cat gpurun_out/mb.log; tail -2 gpurun_out/bench.log
