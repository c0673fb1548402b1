#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench4.log 2> gpurun_out/bench4.err
tail -c 1800 gpurun_out/bench4.log; tail -3 gpurun_out/bench4.err
