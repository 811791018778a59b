"""Phase timeline of CTA 0 of attention_tc4 (CS_ATTN_DBG bit 16): clock64 at the phase boundaries of each role."""
import ctypes, os, sys
os.environ["CS_ATTN_DBG"] = str(16 | int(os.environ.get("CS_ATTN_DBG", "0")))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clipself_b200 import ops, _lib
B, N, H = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 197, 12
D = H * 64
dev = torch.device("cuda")
qkv = torch.randn(B * N, 3 * D, device=dev).to(torch.bfloat16)
out = torch.empty(B * N, D, device=dev, dtype=torch.bfloat16)
stats = torch.empty(B * N, 4 * H, 2, device=dev)
for _ in range(3):
    ops.attention_fwd(qkv, B, N, H, 0.125, out, None, stats)
torch.cuda.synchronize()
buf = np.zeros((20, 256), dtype=np.uint64)
L = _lib.lib()
L.cs_debug_attn_timeline.argtypes = [ctypes.c_void_p]
assert L.cs_debug_attn_timeline(buf.ctypes.data) == 0
t = buf.astype(np.int64)
items = range(int(sys.argv[2]) if len(sys.argv) > 2 else 3, int(sys.argv[3]) if len(sys.argv) > 3 else 8)
t0 = t[0, 6 * items[0]]
ev = []
names = ["S ready", "max done", "exp done", "O ready", "O read", "epi done"]
for w, tag in ((0, "slot0"), (8, "slot1")):
    for i in items:
        for k in range(6):
            ev.append((t[w, 6 * i + k] - t0, f"{tag} item {i}: {names[k]}"))
mn = ["S0 wait", "S0 issue", "S1 wait", "S1 issue", "PV0 wait", "PV0 issue", "PV1 wait", "PV1 issue"]
for i in items:
    for k in range(8):
        ev.append((t[17, 8 * i + k] - t0, f"  mma  item {i}: {mn[k]}"))
    for k, n in enumerate(["K", "Q0", "Q1", "V"]):
        ev.append((t[16, 4 * i + k] - t0, f"    tma item {i}: {n} issued"))
for c, s in sorted(ev):
    print(f"{c:8d}  {s}")
per = (t[0, 6 * items[-1]] - t[0, 6 * items[0]]) / (len(items) - 1)
print(f"period per item: {per:.0f} clk")
