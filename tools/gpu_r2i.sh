#!/bin/bash
# round 2, call I: streaming state I/O -- parity test and the bench line with the streamed e2e leg
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vp_system.py -x -q -k "streaming" 2>&1 | tail -15 > gpurun_out/pytest_stream.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-secondary > gpurun_out/bench.log 2> gpurun_out/bench.err
cat gpurun_out/pytest_stream.log; tail -1 gpurun_out/bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])
"; tail -3 gpurun_out/bench.err
