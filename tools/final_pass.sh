#!/bin/bash
# round-2 final measurement pass: full GPU suite, bench lines of every workload, ncu launch list of one cfg2 step,
# ncu --set full of the attention kernel and the two dominant GEMM instantiations at the step's shape, region kernels
cd "$(dirname "$0")/.."
O=gpurun_out/final_pass; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee $O/pytest_gpu.txt
timeout 500 python bench.py --steps 20 --warmup 3 2>$O/bench_cfg2.err | tail -1 > $O/bench_cfg2.json; cut -c1-300 $O/bench_cfg2.json
for w in cfg4 cfg5 recipe_b16; do timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 > $O/bench_$w.json; cut -c1-200 $O/bench_$w.json; done
CLIPSELF_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python bench.py --workload cfg2 --profile-one-step > $O/ncu_launches.log 2>&1
python tools/ncu_summarize.py $O/launches.csv > $O/launches_summary.txt; head -24 $O/launches_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_tc4 -s 2 -c 1 -o $O/attn_tc4 python tools/attn_one.py 512 197 12 > $O/ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gemm_kernel -s 1 -c 1 -o $O/gemm_swiglu_fold python tools/gemm_one.py swiglu_fold 512 > $O/ncu_gemm1.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gemm_kernel -s 1 -c 1 -o $O/gemm_proj_emit python tools/gemm_one.py proj_emit 512 > $O/ncu_gemm2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes.sum.per_second,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"roi_align|mask_pool|cosine|l2norm|extract_rois|roi_weights" --csv --log-file $O/region_ncu.csv python tools/region_bench.py > $O/region_bench.txt 2>&1
tail -12 $O/region_bench.txt
ls -la $O
