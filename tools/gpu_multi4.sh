#!/bin/bash
# 4 GPUs: two-stream two-part stages -- parity of the pipelined process-grid cases, bench lines for 2x2 and 1x4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -rs -k "iaw_tiles or streams" 2>&1 | tail -8 > gpurun_out/pytest_multi4b.log
for grid in 2x2 1x4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --grid $grid --steps 4 --warmup 3 --no-cpu --no-e2e --no-secondary > gpurun_out/bench4_$grid.log 2> gpurun_out/bench4_$grid.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 3 --warmup 2 --no-cpu --no-e2e --no-secondary --workload streams > gpurun_out/bench4_streams.log 2> gpurun_out/bench4_streams.err
cat gpurun_out/pytest_multi4b.log; for f in gpurun_out/bench4_2x2.log gpurun_out/bench4_1x4.log gpurun_out/bench4_streams.log; do tail -1 $f | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['config']['decomposition'], d['config']['workload'][:40], 'value', d['value'], 'ms', d['ms_per_step'], 'avg', d['roofline']['avg_launch_ms'], 'launches', d['roofline']['timed_launches'], d['clocks'])
"; done; tail -2 gpurun_out/bench4_1x4.err
