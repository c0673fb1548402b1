#!/bin/bash
# 4 GPUs: process-grid parity tests (2 ranks and 4 ranks), a 4-GPU bench line
mkdir -p gpurun_out
rm -f gpurun_out/multi_gpu_parity.jsonl
nvidia-smi -L > gpurun_out/multi4_gpus.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q -rs 2>&1 | tail -25 > gpurun_out/pytest_multi4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench4.log 2> gpurun_out/bench4.err
cat gpurun_out/pytest_multi4.log; tail -1 gpurun_out/bench4.log | cut -c1-1800; tail -3 gpurun_out/bench4.err
