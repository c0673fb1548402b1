#!/bin/bash
# tests + microbench (order 4 and 6) + bench line, no ncu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest.log
rm -f gpurun_out/mb.log
for args in "" "--order 6"; do
  timeout 300 python tools/microbench_rhs.py $args >> gpurun_out/mb.log 2>&1
done
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
cat gpurun_out/pytest.log gpurun_out/mb.log; tail -c 2500 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
