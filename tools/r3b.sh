#!/bin/bash
mkdir -p gpurun_out/r3b
for l in 0 1; do for s in 1 2 3 4 6 8; do ./tools/microbench/tma_stream $l $s 256 | tee -a gpurun_out/r3b/tma_stream.txt; done; done
./tools/microbench/tma_stream 0 8 512 | tee -a gpurun_out/r3b/tma_stream.txt
./tools/microbench/tma_stream 1 8 512 | tee -a gpurun_out/r3b/tma_stream.txt
timeout 600 python -m pytest tests/test_gpu_rowops.py -q -m gpu -k "attention" > gpurun_out/r3b/pytest_attn.txt 2>&1
tail -8 gpurun_out/r3b/pytest_attn.txt
