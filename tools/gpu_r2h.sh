#!/bin/bash
# round 2, call H (1 GPU): tile-subset launches + the whole suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipe.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_pipe.log
timeout 1800 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_pipe.py 2>&1 | tail -15 > gpurun_out/pytest_rest.log
cat gpurun_out/pytest_pipe.log gpurun_out/pytest_rest.log
