#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi2.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench2.log 2> gpurun_out/bench2.err
cat gpurun_out/smi2.log gpurun_out/pytest_multi.log; tail -c 2000 gpurun_out/bench2.log; tail -5 gpurun_out/bench2.err
