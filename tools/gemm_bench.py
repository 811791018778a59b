"""Micro-benchmark of cs_gemm_bf16 over the shapes of the CLIPSelf step (CUDA-event timed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clipself_b200 import ops, _lib as L
from clipself_b200.tower import rope_tables, rope_vectors

dev = torch.device("cuda")
L.require_device()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(name, M, N, K, mode=L.EPI_STORE, out_dtype=torch.bfloat16, residual=False, reps=10):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = torch.randn(N, K, device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    outN = N // 2 if mode == L.EPI_SWIGLU else N
    out = torch.empty(M, outN, device=dev, dtype=out_dtype)
    kw = dict(mode=mode, bias=bias)
    if mode == L.EPI_QKV_ROPE:
        cos, sin = (t.to(dev) for t in rope_vectors(14, 64, 16))
        kw.update(rope=(cos, sin), tokens=197, rope_cols=N // 3 * 2)
    if residual:
        kw.update(residual=out)
    for _ in range(2):
        ops.gemm(a, w, out, M=M, N=N, K=K, **kw)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(a, w, out, M=M, N=N, K=K, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:34s} M={M:6d} N={N:5d} K={K:5d}  {t*1e3:8.1f} us  {2.0*M*N*K/t/1e9:8.1f} TFLOP/s", flush=True)


for imgs in (64, 128, 256):
    M = imgs * 197
    print(f"--- chunk of {imgs} images")
    bench("qkv+rope (bf16 out)", M, 2304, 768, L.EPI_QKV_ROPE)
    bench("proj (+res, f32 out)", M, 768, 768, out_dtype=torch.float32, residual=True)
    bench("w12 swiglu (bf16 out)", M, 4096, 768, L.EPI_SWIGLU)
    bench("w3 (+res, f32 out)", M, 768, 2048, out_dtype=torch.float32, residual=True)
    bench("w12 store (bf16 out, student)", M, 4096, 768)
    bench("plain store bf16 N=768", M, 768, 768)
print("--- wgrad shapes (K = tokens)")
bench("dW3  [768 x 2048], K=12608", 768, 2048, 12608, out_dtype=torch.float32)
bench("dW12 [4096 x 768], K=12608", 4096, 768, 12608, out_dtype=torch.float32)
bench("dWqkv [2304 x 768], K=12608", 2304, 768, 12608, out_dtype=torch.float32)
bench("square 8192^3 (bf16 out)", 8192, 8192, 8192)
