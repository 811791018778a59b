#!/bin/bash
# attention kernels: parity tests, timing at the teacher's shape (with the CS_ATTN_DBG ablations) and the phase timeline
O=gpurun_out/attn_check; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_rowops.py -q -m gpu -x -k attention 2>&1 | tail -3 | tee $O/pytest_attn.txt
grep -q passed $O/pytest_attn.txt && ! grep -q failed $O/pytest_attn.txt || exit 1
for b in 256 512; do timeout 120 python tools/attn_one.py $b 197 12 2>&1 | tail -1 | tee -a $O/timing.txt; done
timeout 120 python tools/attn_timeline.py 256 3 6 > $O/timeline.txt 2>&1; tail -1 $O/timeline.txt
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee $O/pytest_gpu.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | tail -1 > $O/bench_cfg2.json; cut -c1-260 $O/bench_cfg2.json
