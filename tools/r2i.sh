#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2i; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_cabi_tower.py -q -m gpu -x -s 2>&1 | grep -E "C ABI|passed|failed|Error|error|assert" | head -20
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_cabi_tower.py 2>&1 | tail -5
timeout 300 python bench.py --workload cfg2 --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $O/bench_cfg2.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], 'img/s', d['ms_per_step'], 'e2e', d['e2e']['value'], 'devcrops', d['e2e_device_crops']['value'], d['e2e_device_crops']['h2d_bytes_per_step'], 'launches', d['gpu_launches'])"
