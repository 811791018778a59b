#!/bin/bash
# round-2 measurement pass: full GPU suite, bench lines, ncu launch list of one step, ncu full of the dominant GEMM at the
# step's M, ncu dram throughput of the region kernels
cd "$(dirname "$0")/.."
O=gpurun_out/r2k; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee $O/pytest_gpu.txt
timeout 400 python bench.py --steps 10 --warmup 3 2>$O/bench_cfg2.err | tail -1 > $O/bench_cfg2.json; cut -c1-700 $O/bench_cfg2.json
for w in cfg4 cfg5 recipe_b16; do timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 > $O/bench_$w.json; cut -c1-330 $O/bench_$w.json; done
CLIPSELF_NO_GRAPH=1 CLIPSELF_PY_TOWER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python bench.py --workload cfg2 --profile-one-step > $O/ncu_launches.log 2>&1
python tools/ncu_summarize.py $O/launches.csv > $O/launches_summary.txt; head -30 $O/launches_summary.txt
# region kernels: dram bytes per second under ncu
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes.sum.per_second,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"roi_align|mask_pool|cosine|l2norm|extract_rois|roi_weights" --csv --log-file $O/region_ncu.csv python tools/region_bench.py > $O/region_bench.txt 2>&1
tail -22 $O/region_bench.txt
