#!/bin/bash
# 2-GPU validation: NCCL gradient test + the driver's 2-rank bench launch (own arm and reference arm) + smoke
O=gpurun_out/two_gpu_check; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_nccl.py -q -m gpu -x 2>&1 | tail -2 | tee $O/pytest_nccl.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu 2>$O/bench_n2.err | tail -1 > $O/bench_cfg2_n2.json; cut -c1-300 $O/bench_cfg2_n2.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
