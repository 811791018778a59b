"""Run a few launches of one GEMM configuration (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clipself_b200 import ops, _lib as L
from clipself_b200.tower import rope_tables, rope_vectors
which = sys.argv[1]
dev = torch.device("cuda")
M = 256 * 197
cfgs = {"store": (4096, 768, L.EPI_STORE, torch.bfloat16, False), "proj": (768, 768, L.EPI_STORE, torch.float32, True),
        "swiglu": (4096, 768, L.EPI_SWIGLU, torch.bfloat16, False), "qkv": (2304, 768, L.EPI_QKV_ROPE, torch.bfloat16, False)}
N, K, mode, odt, res = cfgs[which]
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = torch.randn(N, K, device=dev).to(torch.bfloat16)
bias = torch.randn(N, device=dev)
out = torch.zeros(M, N // 2 if mode == L.EPI_SWIGLU else N, device=dev, dtype=odt)
kw = dict(mode=mode, bias=bias)
if mode == L.EPI_QKV_ROPE:
    cos, sin = (t.to(dev) for t in rope_vectors(14, 64, 16))
    kw.update(rope=(cos, sin), tokens=197, rope_cols=1536)
if res:
    kw.update(residual=out)
for _ in range(3):
    ops.gemm(a, w, out, M=M, N=N, K=K, **kw)
torch.cuda.synchronize()
