#!/bin/bash
# round-2 session-2: attention_tc4 parity + timing against attention_tc3
mkdir -p gpurun_out/r3i2
timeout 600 python -m pytest tests/test_gpu_rowops.py -q -m gpu -k "attention" -x > gpurun_out/r3i2/pytest_attn.txt 2>&1
tail -5 gpurun_out/r3i2/pytest_attn.txt
for b in 256 512; do
  timeout 120 python tools/attn_one.py $b 197 12 2>&1 | tail -1 | sed "s/^/tc4 /" | tee -a gpurun_out/r3i2/timing.txt
  CS_ATTN_V3=1 timeout 120 python tools/attn_one.py $b 197 12 2>&1 | tail -1 | sed "s/^/tc3 /" | tee -a gpurun_out/r3i2/timing.txt
done
for d in 1 2 4 8 3 7 15; do
  CS_ATTN_DBG=$d timeout 120 python tools/attn_one.py 256 197 12 2>&1 | tail -1 | sed "s/^/tc4 dbg=$d /" | tee -a gpurun_out/r3i2/timing.txt
done
python tools/attn_timeline.py 256 3 6 > gpurun_out/r3i2/timeline.txt 2>&1; tail -2 gpurun_out/r3i2/timeline.txt
