#!/bin/bash
O=gpurun_out/r3u; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_rowops.py -q -m gpu -x -k attention 2>&1 | tail -4 | tee $O/pytest_attn.txt
grep -q passed $O/pytest_attn.txt && ! grep -q failed $O/pytest_attn.txt || exit 1
timeout 100 python tools/attn_one.py 256 577 16 | tail -1 | tee -a $O/timing.txt
timeout 100 python tools/attn_one.py 2 4097 12 | tail -1 | tee -a $O/timing.txt
timeout 100 python tools/attn_one.py 64 257 16 | tail -1 | tee -a $O/timing.txt
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee $O/pytest_gpu.txt
for w in cfg4 recipe_b16; do timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 > $O/bench_$w.json; cut -c1-200 $O/bench_$w.json; done
