#!/bin/bash
O=gpurun_out/r3v; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_rowops.py tests/test_gpu_cabi_tower.py -q -m gpu -x -k "attention_cls or cabi or tower" 2>&1 | tail -5 | tee $O/pytest_cls.txt
grep -q passed $O/pytest_cls.txt && ! grep -q failed $O/pytest_cls.txt || exit 1
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee $O/pytest_gpu.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu 2>$O/bench_cfg2.err | tail -1 > $O/bench_cfg2.json; cut -c1-330 $O/bench_cfg2.json
CLIPSELF_FULL_LAST_BLOCK=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | tail -1 > $O/bench_cfg2_full_last_block.json; cut -c1-330 $O/bench_cfg2_full_last_block.json
