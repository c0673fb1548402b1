#!/bin/bash
# round 2, call G: deck options (Krook / JB / open boundaries) + whole suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vp_system.py -x -q -k "deck_options" 2>&1 | tail -30 > gpurun_out/pytest_new.log
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_all.log
cat gpurun_out/pytest_new.log gpurun_out/pytest_all.log
