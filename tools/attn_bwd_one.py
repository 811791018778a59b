"""A few launches of cs_attention_bwd (for ncu) + CUDA-event timing.  argv: B N H [rope 0|1]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clipself_b200 import ops
from clipself_b200.tower import rope_tables
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 197
H = int(sys.argv[3]) if len(sys.argv) > 3 else 12
D = H * 64
g = int(round((N - 1) ** 0.5))
rope = tuple(t.to(dev) for t in rope_tables(g, 64, 16)) if g * g == N - 1 else None
qkv = (torch.randn(B * N, 3 * D, device=dev) * 0.7).to(torch.bfloat16)
out = torch.empty(B * N, D, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B * H * N, device=dev)
ops.attention_fwd(qkv, B, N, H, 0.125, out, lse)
d_out = torch.randn(B * N, D, device=dev).to(torch.bfloat16)
dqkv = torch.empty(B * N, 3 * D, device=dev, dtype=torch.bfloat16)
delta = torch.empty(B * H * N, device=dev)
for _ in range(2):
    ops.attention_bwd(qkv, out, d_out, lse, B, N, H, 0.125, rope, delta, dqkv)
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.attention_bwd(qkv, out, d_out, lse, B, N, H, 0.125, rope, delta, dqkv); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = sorted(ts)[2]
print(f"attention bwd B={B} N={N} H={H} legacy={os.environ.get('CS_ATTN_LEGACY', '0')}: {t*1e3:.1f} us, "
      f"{10.0*B*H*N*N*64/t/1e9:.1f} TFLOP/s (useful, 5 contractions)")
