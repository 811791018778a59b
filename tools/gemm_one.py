"""Run a few launches of one GEMM configuration (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clipself_b200 import ops, _lib as L
from clipself_b200.tower import rope_tables, rope_vectors
which = sys.argv[1]
dev = torch.device("cuda")
M = (int(sys.argv[2]) if len(sys.argv) > 2 else 256) * 197      # crops x tokens (the step runs 512-crop chunks)
cfgs = {"store": (4096, 768, L.EPI_STORE, torch.bfloat16, False), "proj": (768, 768, L.EPI_STORE, torch.float32, True),
        "swiglu": (4096, 768, L.EPI_SWIGLU, torch.bfloat16, False), "qkv": (2304, 768, L.EPI_QKV_ROPE, torch.bfloat16, False),
        # the teacher's folded launches: norm2 folded into w1|w2 + SiLU*mul + row statistics (the dominant kernel of the step)
        "swiglu_fold": (4096, 768, L.EPI_SWIGLU, torch.bfloat16, False), "proj_emit": (768, 768, L.EPI_STORE, torch.float32, True)}
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
if which in ("swiglu_fold", "proj_emit"):
    kw.update(ln_fold=(torch.rand(M, 6, 2, device=dev) + 1.0, torch.randn(N, device=dev), 6, K, 1e-6))
if which == "swiglu_fold":
    kw.update(stats_out=torch.empty(M, N // 128, 2, device=dev))
if which == "proj_emit":
    kw.update(out2=torch.empty(M, N, device=dev, dtype=torch.bfloat16), stats_out=torch.empty(M, 6, 2, device=dev))
for _ in range(3):
    ops.gemm(a, w, out, M=M, N=N, K=K, **kw)
torch.cuda.synchronize()
