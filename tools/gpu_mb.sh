#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/mb.log
for args in "" "--mode rhs" "128 128 128 128" "$@"; do
  timeout 300 python tools/microbench_rhs.py $args >> gpurun_out/mb.log 2>&1
done
cat gpurun_out/mb.log
