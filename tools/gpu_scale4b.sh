#!/bin/bash
# one 4-GPU box: round-1 kernel (LK_NO_PIPE) at N = 1 and 4, and the y-only grid 1x4 with the pipelined kernel
mkdir -p gpurun_out
run() { # name, nproc, extra env, extra args
  if [ "$2" = "1" ]; then
    env $3 timeout 600 python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu --no-e2e --no-secondary $4 > gpurun_out/scale_$1.log 2> gpurun_out/scale_$1.err
  else
    env $3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $2 --steps 4 --warmup 3 --no-cpu --no-e2e --no-secondary $4 > gpurun_out/scale_$1.log 2> gpurun_out/scale_$1.err
  fi
  tail -1 gpurun_out/scale_$1.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('$1', d['config']['decomposition'], 'value', round(d['value']/1e9,1), 'ms', round(d['ms_per_step'],2), 'kernel', round(r['avg_launch_ms'],2), 'share', r['kernel_share_of_step'], d['clocks']['sm_mhz'])
"
}
run n1_march 1 LK_NO_PIPE=1
run n4_march 4 LK_NO_PIPE=1
run n1 1 A=1
run n4_1x4 4 A=1 "--grid 1x4"
run n4_2x2 4 A=1
