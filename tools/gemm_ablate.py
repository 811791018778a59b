"""Epilogue ablation of cs_gemm_bf16: dbg=8 skips the epilogue body (mainloop + handshakes only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clipself_b200 import ops, _lib as L
from clipself_b200.tower import rope_vectors
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
M = int(os.environ.get("M_CROPS", "128")) * 197


def run(name, N, K, mode, odt, res, dbg, fold=False, emit=False):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = torch.randn(N, K, device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    out = torch.zeros(M, N // 2 if mode == L.EPI_SWIGLU else N, device=dev, dtype=odt)
    kw = dict(mode=mode, bias=bias, dbg=dbg)
    if mode == L.EPI_QKV_ROPE:
        kw.update(rope=tuple(t.to(dev) for t in rope_vectors(14, 64, 16)), tokens=197, rope_cols=N // 3 * 2)
    if res == 1:
        kw.update(residual=torch.zeros_like(out))
    if res == 2:
        kw.update(residual=out)
    if fold:
        kw.update(ln_fold=(torch.rand(M, 6, 2, device=dev) + 1.0, torch.randn(N, device=dev), 6, K, 1e-6))
    if emit:
        kw.update(residual=out, out2=torch.empty(M, N, device=dev, dtype=torch.bfloat16),
                  stats_out=torch.empty(M, 2 * ((N + 255) // 256), 2, device=dev))
    if mode == L.EPI_SWIGLU:
        kw.update(stats_out=torch.empty(M, N // 128, 2, device=dev))
    ts = []
    for i in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(a, w, out, M=M, N=N, K=K, **kw); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:28s} dbg={dbg:2d} {t*1e3:8.1f} us {2.0*M*N*K/t/1e9:8.1f} TFLOP/s", flush=True)


import sys as _s
if len(_s.argv) > 1 and _s.argv[1] == "bits":
    for dbg in (0, 1, 2, 4, 5, 6, 8):
        run("w12 store bf16 N=4096", 4096, 768, L.EPI_STORE, torch.bfloat16, 0, dbg)
    _s.exit(0)
if len(_s.argv) > 1 and _s.argv[1] == "emit":
    for dbg in (0, 16, 1, 8):          # 16: no L2 prefetch of the next tile's residual; 1: no global stores; 8: mainloop only
        run("proj fold emit N=768", 768, 768, L.EPI_STORE, torch.float32, 0, dbg, fold=True, emit=True)
        run("w3 fold emit N=768 K=2048", 768, 2048, L.EPI_STORE, torch.float32, 0, dbg, fold=True, emit=True)
    _s.exit(0)
for dbg in (0, 8):
    run("w12 store bf16 N=4096", 4096, 768, L.EPI_STORE, torch.bfloat16, 0, dbg)
    run("qkv rope bf16 N=2304", 2304, 768, L.EPI_QKV_ROPE, torch.bfloat16, 0, dbg)
    run("swiglu N=4096", 4096, 768, L.EPI_SWIGLU, torch.bfloat16, 0, dbg)
    run("proj red f32 N=768", 768, 768, L.EPI_STORE, torch.float32, 2, dbg)
    run("proj res-load f32 N=768", 768, 768, L.EPI_STORE, torch.float32, 1, dbg)
    run("w3 red f32 N=768 K=2048", 768, 2048, L.EPI_STORE, torch.float32, 2, dbg)
    run("store f32 N=768 K=768", 768, 768, L.EPI_STORE, torch.float32, 0, dbg)
    run("qkv rope fold N=2304", 2304, 768, L.EPI_QKV_ROPE, torch.bfloat16, 0, dbg, fold=True)
    run("swiglu fold N=4096", 4096, 768, L.EPI_SWIGLU, torch.bfloat16, 0, dbg, fold=True)
    run("proj fold emit N=768", 768, 768, L.EPI_STORE, torch.float32, 0, dbg, fold=True, emit=True)
    run("w3 fold emit N=768 K=2048", 768, 2048, L.EPI_STORE, torch.float32, 0, dbg, fold=True, emit=True)
