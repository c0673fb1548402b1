#!/bin/bash
# round 2, call F: flux diagnostics + fold -- new tests first, then the whole suite and a bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flux.py tests/test_gpu_f77abi.py tests/test_gpu_pipe.py -x -q 2>&1 | tail -25 > gpurun_out/pytest_new.log
timeout 600 python -m pytest tests/test_gpu_vp_system.py -x -q -k "flux_histories" 2>&1 | tail -25 >> gpurun_out/pytest_new.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_pipe.py --deselect tests/test_gpu_flux.py --deselect tests/test_gpu_f77abi.py 2>&1 | tail -15 > gpurun_out/pytest_rest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.log 2> gpurun_out/bench.err
cat gpurun_out/pytest_new.log gpurun_out/pytest_rest.log; tail -1 gpurun_out/bench.log | cut -c1-1500; tail -3 gpurun_out/bench.err
