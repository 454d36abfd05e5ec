#!/bin/bash
# Run each GPU test function in its own process (a trapped kernel poisons the CUDA context),
# each under its own timeout; collect logs into gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/probe_gpu.txt 2>&1
FILES=${@:-tests/test_kernels_gpu.py}
: > gpurun_out/probe_summary.txt
for f in $FILES; do
  for t in $(python -m pytest $f --collect-only -q -m gpu 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u); do
    name=$(echo $t | sed 's/.*:://')
    timeout 300 python -m pytest "$t" -q -m gpu -s --no-header -p no:cacheprovider > gpurun_out/probe_$name.log 2>&1
    rc=$?
    echo "$name rc=$rc $(tail -1 gpurun_out/probe_$name.log)" >> gpurun_out/probe_summary.txt
  done
done
cat gpurun_out/probe_summary.txt
