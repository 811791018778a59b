#!/bin/bash
O=gpurun_out/r3j; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee $O/pytest_gpu.txt
timeout 400 python bench.py --steps 10 --warmup 3 2>$O/bench_cfg2.err | tail -1 > $O/bench_cfg2.json; cut -c1-400 $O/bench_cfg2.json
