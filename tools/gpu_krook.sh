#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipe.py -x -q -k "krook" 2>&1 | tail -12 > gpurun_out/krook.log
timeout 600 python -m pytest tests/test_gpu_vp_system.py -x -q -k "deck_options" 2>&1 | tail -8 >> gpurun_out/krook.log
cat gpurun_out/krook.log
