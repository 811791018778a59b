#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2c; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x 2>&1 | tail -15
timeout 200 python tools/gemm_ablate.py 2>&1 | tee $O/gemm_ablate.txt
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_gemm.py 2>&1 | tail -15
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $O/bench_cfg2.json | cut -c1-2000
CLIPSELF_NO_NORM_FOLD=1 timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $O/bench_cfg2_nonormfold.json | cut -c1-400
timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $O/bench_cfg4.json | cut -c1-400
