#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2b; mkdir -p $O
for shape in "64 197 12" "16 577 16" "2 4097 12"; do
  env -u CS_ATTN_BWD_TC timeout 120 python tools/attn_bwd_one.py $shape 2>&1 | tail -1
  CS_ATTN_BWD_TC=1 timeout 120 python tools/attn_bwd_one.py $shape 2>&1 | tail -1
done
CS_ATTN_LONG_TC=1 timeout 120 python tools/attn_one.py 2 4097 12 | tail -1
env -u CS_ATTN_LONG_TC timeout 120 python tools/attn_one.py 2 4097 12 | tail -1
echo "--- bench A/B"
for w in cfg4 recipe_b16; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330
  CS_ATTN_LONG_TC=1 CS_ATTN_BWD_TC=1 timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330
done
CS_ATTN_BWD_TC=1 timeout 300 python bench.py --workload cfg2 --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330
echo "--- ncu"
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:attention_fwd_tc_kernel -s 2 -c 1 -o $O/attn_short python tools/attn_one.py 256 197 12 > $O/ncu_attn_short.log 2>&1
CS_ATTN_LONG_TC=1 timeout 300 $NCU -k regex:attention_fwd_tc_long -s 2 -c 1 -o $O/attn_long python tools/attn_one.py 64 577 16 > $O/ncu_attn_long.log 2>&1
timeout 300 $NCU -k regex:gemm_kernel -s 1 -c 1 -o $O/gemm_qkv python tools/gemm_one.py qkv > $O/ncu_gemm_qkv.log 2>&1
timeout 300 $NCU -k regex:gemm_kernel -s 1 -c 1 -o $O/gemm_proj python tools/gemm_one.py proj > $O/ncu_gemm_proj.log 2>&1
ls -la $O
