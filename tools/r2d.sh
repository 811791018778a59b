#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2d; mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x -k "store_f32 or residual_inplace" 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x 2>&1 | tail -12
timeout 200 python tools/gemm_ablate.py 2>&1 | tee $O/gemm_ablate_2cta.txt
CS_GEMM_1CTA=1 timeout 200 python tools/gemm_ablate.py 2>&1 | head -11 | tee $O/gemm_ablate_1cta.txt
