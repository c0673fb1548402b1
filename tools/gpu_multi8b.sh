#!/bin/bash
# 8 GPUs: the contract line with the y-only grid (secondary = streams on its 2x4 grid)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu > gpurun_out/bench8b.log 2> gpurun_out/bench8b.err
tail -1 gpurun_out/bench8b.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['config']['decomposition'], d['config']['workload'][:60], 'value', d['value'], 'ms', d['ms_per_step'], 'avg', d['roofline']['avg_launch_ms'], 'share', d['roofline']['kernel_share_of_step'], d['clocks'])
print('e2e', d['e2e'])
sec = d['config'].get('secondary')
if sec: print('secondary', sec['config']['decomposition'], sec['config']['workload'][:60], sec['value'], sec['ms_per_step'], sec['roofline']['avg_launch_ms'])
"; tail -3 gpurun_out/bench8b.err
