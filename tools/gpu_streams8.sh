#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu --no-e2e --workload streams > gpurun_out/bench_streams8.log 2> gpurun_out/bench_streams8.err
tail -c 2200 gpurun_out/bench_streams8.log; tail -5 gpurun_out/bench_streams8.err; nvidia-smi --query-gpu=memory.used --format=csv | head -3
