#!/bin/bash
# round 2, call D: velocity-boundary fold -- pipe parity, system tests, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipe.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_pipe.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_pipe.py 2>&1 | tail -15 > gpurun_out/pytest_rest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.log 2> gpurun_out/bench.err
cat gpurun_out/pytest_pipe.log gpurun_out/pytest_rest.log; tail -1 gpurun_out/bench.log | cut -c1-1500; tail -3 gpurun_out/bench.err
