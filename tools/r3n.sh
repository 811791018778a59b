#!/bin/bash
O=gpurun_out/r3n; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_rowops.py tests/test_gpu_backward_kernels.py -q -m gpu -x 2>&1 | tail -3 | tee $O/pytest_attn.txt
timeout 100 python tools/attn_one.py 256 577 16 | tail -1 | tee -a $O/timing.txt
timeout 100 python tools/attn_one.py 2 4097 12 | tail -1 | tee -a $O/timing.txt
timeout 100 python tools/attn_bwd_one.py 64 197 12 | tail -1 | tee -a $O/timing.txt
timeout 100 python tools/attn_bwd_one.py 16 577 16 | tail -1 | tee -a $O/timing.txt
timeout 200 python tools/crops_time.py 2>&1 | tail -4 | tee $O/crops_time.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/crops_launches.csv python tools/crops_time.py > /dev/null 2>&1
python tools/ncu_summarize.py $O/crops_launches.csv 2>/dev/null | head -14 | tee $O/crops_launches_summary.txt
