#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2h; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_region.py tests/test_gpu_step.py -q -m gpu -x -k "raw_image or cli_end_to_end or eval_after or device_crops" 2>&1 | tail -15
timeout 300 python bench.py --workload cfg2 --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $O/bench_cfg2.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], 'img/s e2e', d['e2e'], 'devcrops', d['e2e_device_crops'])"
