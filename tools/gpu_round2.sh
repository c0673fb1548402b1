#!/bin/bash
# one gpurun call: parity tests, bench line (own + reference arm), ncu launch list, DRAM traffic per launch, one
# full capture of the stage kernel.  Outputs under gpurun_out/; tools/ncu_*.py summarise them into profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_stage_ --csv --log-file gpurun_out/traffic.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/bench_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_pipe -s 1 -c 1 -o gpurun_out/prof_pipe -f \
  python tools/microbench_rhs.py 256 256 128 128 --reps 1 --fold > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest.log; tail -1 gpurun_out/bench.log | cut -c1-600; tail -2 gpurun_out/bench.err; tail -1 gpurun_out/bench_ref.log | cut -c1-300
