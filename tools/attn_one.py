"""A few launches of cs_attention_fwd at the teacher's shape (for ncu) + CUDA-event timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clipself_b200 import ops
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 197
H = int(sys.argv[3]) if len(sys.argv) > 3 else 12
D = H * 64
qkv = torch.randn(B * N, 3 * D, device=dev).to(torch.bfloat16)
out = torch.empty(B * N, D, device=dev, dtype=torch.bfloat16)
stats = torch.empty(B * N, 4 * H, 2, device=dev)
for _ in range(2):
    ops.attention_fwd(qkv, B, N, H, 0.125, out, None, stats)
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.attention_fwd(qkv, B, N, H, 0.125, out, None, stats); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = sorted(ts)[2]
print(f"attention fwd B={B} N={N} H={H}: {t*1e3:.1f} us, {4.0*B*H*N*N*64/t/1e9:.1f} TFLOP/s (useful)")
