#!/bin/bash
# order-6 iteration: parity tests that involve order 6 + microbenchmarks
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "6 or order or weno or rhs or step" 2>&1 | tail -8 > gpurun_out/pytest_o6.log
rm -f gpurun_out/mb_o6.log
for args in "--order 6" "--order 6 --mode rhs" ""; do
  timeout 300 python tools/microbench_rhs.py $args >> gpurun_out/mb_o6.log 2>&1
done
cat gpurun_out/pytest_o6.log gpurun_out/mb_o6.log
