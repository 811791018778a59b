#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_rowops.py -q -m gpu -x -k attention -s 2>&1 | tail -12
echo "--- timing v2 vs v1"
timeout 120 python tools/attn_one.py 256 197 12 | tail -1
CS_ATTN_V1=1 timeout 120 python tools/attn_one.py 256 197 12 | tail -1
timeout 120 python tools/attn_one.py 512 197 12 | tail -1; timeout 120 python tools/attn_one.py 64 197 12 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity_stages.py tests/test_gpu_backward_kernels.py -q -m gpu -x 2>&1 | tail -4
