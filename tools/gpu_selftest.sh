#!/bin/bash
# Run every GPU test file in its own process (a trap in one kernel must not poison the rest),
# logs under gpurun_out/selftest/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/selftest
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/selftest/gpu.txt 2>&1
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 600 python -m pytest "$f" -q -s -m gpu -p no:cacheprovider > "gpurun_out/selftest/$name.log" 2>&1
  echo "$name exit=$?" | tee -a gpurun_out/selftest/summary.txt
  tail -n 3 "gpurun_out/selftest/$name.log"
done
