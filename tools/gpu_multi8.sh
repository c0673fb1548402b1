#!/bin/bash
# 8 GPUs: 2x4 / 4x2 process-grid parity tests, the contract's bench line (secondary = streams), streams with and without two-part stages
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi8_gpus.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -rs -k "test_process_grids and 8-" 2>&1 | tail -8 > gpurun_out/pytest_multi8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu > gpurun_out/bench8.log 2> gpurun_out/bench8.err
LOKI_SPLIT_STAGES=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu --no-e2e --no-secondary --workload streams > gpurun_out/bench8_streams_nosplit.log 2> gpurun_out/bench8_streams_nosplit.err
cat gpurun_out/pytest_multi8.log; for f in gpurun_out/bench8.log gpurun_out/bench8_streams_nosplit.log; do tail -1 $f | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['config']['decomposition'], d['config']['workload'][:50], 'value', d['value'], 'ms', d['ms_per_step'], 'avg', d['roofline']['avg_launch_ms'], 'share', d['roofline']['kernel_share_of_step'], d['clocks'])
sec = d['config'].get('secondary')
if sec: print('secondary', sec['config']['workload'][:60], sec['value'], sec['ms_per_step'], sec['roofline']['avg_launch_ms'], sec['roofline'].get('kernel_share_of_step'), sec['roofline'].get('launches_note'))
"; done; tail -2 gpurun_out/bench8.err
