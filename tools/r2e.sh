#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2e; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $O/bench_cfg2.json | cut -c1-1800
CLIPSELF_NO_NORM_FOLD=1 timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $O/bench_cfg2_nonormfold.json | cut -c1-330
CS_GEMM_1CTA=1 timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $O/bench_cfg2_1cta.json | cut -c1-330
timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $O/bench_cfg4.json | cut -c1-330
