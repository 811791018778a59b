#!/bin/bash
O=gpurun_out/r3t; mkdir -p $O
for c in 96 128 192 296 512; do
  CLIPSELF_TEACHER_CHUNK=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | tail -1 > $O/bench_chunk$c.json
  python - <<PY
import json
d=json.load(open("$O/bench_chunk$c.json"))
print("chunk $c:", d["value"], "img/s", d["ms_per_step"], "ms; e2e", d["e2e"]["value"], "; gemm", d["roofline"]["achieved"], d["roofline"]["gemm_share_of_step"], d["clocks"]["sm_mhz"])
PY
done
