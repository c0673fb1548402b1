#!/bin/bash
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/bench8_full.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench8_full.log 2>> gpurun_out/bench8_full.err
tail -c 2500 gpurun_out/bench8_full.log; tail -5 gpurun_out/bench8_full.err
