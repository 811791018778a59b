#!/bin/bash
# Round-2 starting point: validate the experimental long-sequence tcgen05 attention forward on a B200
# (own process, bounded by `timeout`; the kernel's mbarrier waits trap instead of hanging).
#   gpurun --timeout 600 -- 'bash tools/attn_long_try.sh'
cd "$(dirname "$0")/.."
CS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_rowops.py -q -m gpu -k long_tc -x -s 2>&1 | tail -15
echo "--- timing (B=256 crops, ViT-L heads) mma.sync vs tcgen05"
env -u CS_ATTN_LONG_TC timeout 120 python tools/attn_one.py 256 577 16 2>&1 | tail -1
CS_ATTN_LONG_TC=1 timeout 120 python tools/attn_one.py 256 577 16 2>&1 | tail -1
echo "--- the teacher's shape (256 crops x 197 tokens x 12 heads): single-pass kernel vs the ping-pong long kernel"
env -u CS_ATTN_LONG_TC -u CS_ATTN_FORCE_LONG timeout 120 python tools/attn_one.py 256 197 12 2>&1 | tail -1
CS_ATTN_LONG_TC=1 CS_ATTN_FORCE_LONG=1 timeout 120 python tools/attn_one.py 256 197 12 2>&1 | tail -1
echo "--- staged tcgen05 backward (never run before): parity of dq / dk / dv against fp32 autograd"
CS_ATTN_BWD_TC=1 timeout 300 python -m pytest tests/test_gpu_backward_kernels.py -q -m gpu -k attention_bwd -x -s 2>&1 | tail -15
echo "--- eval loop / shared dense pass (added after the round-1 GPU budget ended)"
CS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_step.py -q -m gpu -k eval_loop -x -s 2>&1 | tail -8
echo "--- staged on-device crop generation (bit-exact vs the reference-transform fixtures)"
CS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_region.py -q -m gpu -k device_crops -x -s 2>&1 | tail -8
