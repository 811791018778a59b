#!/bin/bash
O=gpurun_out/r3o; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x 2>&1 | tail -5 | tee $O/pytest_gemm.txt
grep -q passed $O/pytest_gemm.txt && ! grep -q failed $O/pytest_gemm.txt || exit 1
true
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee $O/pytest_gpu.txt
timeout 400 python bench.py --steps 10 --warmup 3 2>$O/bench_cfg2.err | tail -1 > $O/bench_cfg2.json; cut -c1-330 $O/bench_cfg2.json
