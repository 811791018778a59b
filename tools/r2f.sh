#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2f; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_tower.py tests/test_gpu_step.py -q -m gpu -x 2>&1 | tail -4
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload cfg2 --steps 6 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print(d['value'], 'img/s', d['ms_per_step'], 'ms e2e', d['e2e']['value'], 'gemm', d['roofline']['achieved'], 'share', d['roofline']['gemm_share_of_step'], 'launches', d['gpu_launches'], 'clk', d['clocks']['sm_mhz'])
except Exception as e: print('ERR', e)
"; }
run CLIPSELF_TEACHER_CHUNK=512
run CLIPSELF_TEACHER_CHUNK=512 CLIPSELF_NO_GRAPH=1
run CLIPSELF_TEACHER_CHUNK=256
run CLIPSELF_TEACHER_CHUNK=128
run CLIPSELF_TEACHER_CHUNK=96
run CLIPSELF_TEACHER_CHUNK=64
run CLIPSELF_TEACHER_CHUNK=128 CLIPSELF_NO_NORM_FOLD=1
run CLIPSELF_TEACHER_CHUNK=64 CLIPSELF_NO_NORM_FOLD=1
