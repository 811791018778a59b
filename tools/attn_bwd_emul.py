"""Numpy emulation of the DATAFLOW of csrc/attention_bwd_tc.cu (staged tcgen05 attention backward), checked against
torch autograd on the CPU.  It follows the kernel's own parameterisation — the same resident / streamed operand
table (r1_col, r2_col, x1_col, x2_col), the same 128-row tiles and 64-row blocks, the same masks, the same bf16
rounding points (P^T / dS tiles), the same accumulator-to-output mapping (ACC1 / ACC2 -> q | k | v sections) and the
same inverse-RoPE convention — so a slip in that bookkeeping (swapped accumulators, wrong section offset, mask on
the wrong index, missing scale) shows up here, without a GPU.  The tensor-core products are plain matmuls.

    python tools/attn_bwd_emul.py
"""
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HD, BM, BS = 64, 128, 64
LOG2E = 1.4426950408889634


def bf16(x):
    return torch.from_numpy(np.asarray(x, np.float32)).to(torch.bfloat16).float().numpy()


def unrope_pair(d0, d1, cs, sn, d):
    """attention_bwd_tc.cu unrope_pair: forward was y0 = x0 c0 - x1 s0, y1 = x1 c1 + x0 s1."""
    c0, c1, s0, s1 = cs[d], cs[d + 1], sn[d], sn[d + 1]
    return d0 * c0 + d1 * s1, d1 * c1 - d0 * s0


def run_pass(dkv, qkv, d_out, lse, delta, B, N, H, scale, rope, dqkv):
    """One launch of attention_bwd_tc_kernel<kDKV>: qkv [B*N,3D] bf16-valued f32, d_out [B*N,D]."""
    D = H * HD
    sl2 = scale * LOG2E
    # operand table exactly as attention_bwd_tc() sets it: (tensor, column offset of head 0)
    if not dkv:
        r1, r2, x1, x2 = (qkv, 0), (d_out, 0), (qkv, D), (qkv, 2 * D)          # Q | dO resident, K | V streamed
    else:
        r1, r2, x1, x2 = (qkv, D), (qkv, 2 * D), (qkv, 0), (d_out, 0)          # K | V resident, Q | dO streamed
    ntiles, nblk = -(-N // BM), -(-N // BS)
    total_rows = B * N

    def tile(op, h, row0, rows):
        t, col = op
        out = np.zeros((rows, HD), np.float32)                                  # TMA: out-of-bounds rows are zero
        lo, hi = row0, min(row0 + rows, total_rows)                             # rows past the image are the NEXT image's
        if hi > lo:
            out[:hi - lo] = t[lo:hi, col + h * HD: col + (h + 1) * HD]
        return out

    for b in range(B):
        for h in range(H):
            for tl in range(ntiles):
                R1, R2 = tile(r1, h, b * N + tl * BM, BM), tile(r2, h, b * N + tl * BM, BM)
                rows = tl * BM + np.arange(BM)                                  # query (dQ) or key (dKV) index per lane
                acc1 = np.zeros((BM, HD), np.float32)
                acc2 = np.zeros((BM, HD), np.float32)
                lse2_r = np.where(rows < N, lse[b, h, np.minimum(rows, N - 1)] * LOG2E, 0.0)
                dl_r = np.where(rows < N, delta[b, h, np.minimum(rows, N - 1)], 0.0)
                for j in range(nblk):
                    X1, X2 = tile(x1, h, b * N + j * BS, BS), tile(x2, h, b * N + j * BS, BS)
                    T1, T2 = R1 @ X1.T, R2 @ X2.T                               # [128 lanes, 64 streamed]
                    cols = j * BS + np.arange(BS)
                    valid = (cols < N)[None, :]
                    if dkv:                                                     # per-COLUMN scalars (the staged vectors)
                        l2 = np.where(cols < N, lse[b, h, np.minimum(cols, N - 1)] * LOG2E, 0.0)[None, :]
                        dl = np.where(cols < N, delta[b, h, np.minimum(cols, N - 1)], 0.0)[None, :]
                    else:                                                       # per-ROW scalars
                        l2, dl = lse2_r[:, None], dl_r[:, None]
                    with np.errstate(over="ignore", invalid="ignore"):
                        p = np.where(valid, np.exp2(T1 * sl2 - l2), 0.0).astype(np.float32)
                        ds = (p * (T2 - dl) * scale).astype(np.float32)
                    if not dkv:
                        acc1 += bf16(ds) @ X1                                   # dQ += dS K
                    else:
                        acc1 += bf16(p) @ X2                                    # dV += P^T dO
                        acc2 += bf16(ds) @ X1                                   # dK += dS^T Q
                rot = acc2 if dkv else acc1                                     # the rotated gradient: dK or dQ
                if rope is not None:
                    cos, sin = rope
                    for r in range(BM):
                        if rows[r] < N and rows[r] > 0:                         # token 0 (CLS) is not rotated
                            for i in range(0, HD, 2):
                                rot[r, i], rot[r, i + 1] = unrope_pair(rot[r, i], rot[r, i + 1], cos[rows[r] - 1], sin[rows[r] - 1], i)
                for r in range(BM):
                    if rows[r] < N:
                        o = (b * N + rows[r])
                        if not dkv:
                            dqkv[o, h * HD:(h + 1) * HD] = acc1[r]                          # q section
                        else:
                            dqkv[o, D + h * HD: D + (h + 1) * HD] = acc2[r]                 # k section
                            dqkv[o, 2 * D + h * HD: 2 * D + (h + 1) * HD] = acc1[r]         # v section


def check(B, N, H, rope, seed=0):
    from clipself_b200.tower import rope_tables
    D = H * HD
    torch.manual_seed(seed)
    raw = (torch.randn(B, N, 3, H, HD) * 0.7).requires_grad_(True)
    cos = sin = None
    if rope:
        g = int(round((N - 1) ** 0.5))
        assert g * g == N - 1
        cos, sin = rope_tables(g, HD, 16)

        def rot(t):
            pairs = t.reshape(*t.shape[:-1], 32, 2)
            r = torch.stack((-pairs[..., 1], pairs[..., 0]), -1).reshape(t.shape)
            return t * cos[None, :, None, None, :] + r * sin[None, :, None, None, :]
        x = torch.cat([torch.cat([raw[:, :1, :2], rot(raw[:, 1:, :2])], dim=1), raw[:, :, 2:]], dim=2)
    else:
        x = raw
    xb = x.to(torch.bfloat16)
    q, k, v = (t.permute(0, 2, 1, 3).float() for t in xb.unbind(2))
    s = (q @ k.transpose(-1, -2)) * 0.125
    out = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B * N, D)
    d_out = torch.randn(B * N, D).to(torch.bfloat16).float()
    out.backward(d_out)
    ref = raw.grad.reshape(B * N, 3 * D).numpy()

    qkv = xb.detach().float().reshape(B * N, 3 * D).numpy()
    lse = torch.logsumexp(s, -1).detach().numpy()                               # [B,H,N], natural log
    o_bf = out.detach().to(torch.bfloat16).float()
    delta = (o_bf * d_out).view(B, N, H, HD).sum(-1).permute(0, 2, 1).numpy()  # attn_delta_kernel
    dqkv = np.full((B * N, 3 * D), np.nan, np.float32)
    rp = (cos.numpy(), sin.numpy()) if rope else None
    run_pass(False, qkv, d_out.numpy(), lse, delta, B, N, H, 0.125, rp, dqkv)
    run_pass(True, qkv, d_out.numpy(), lse, delta, B, N, H, 0.125, rp, dqkv)
    assert not np.isnan(dqkv).any(), "some gradient rows were never written"
    res = {}
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        a, b_ = bf16(dqkv[:, sl]), ref[:, sl]
        res[name] = float(np.linalg.norm(a - b_) / np.linalg.norm(b_))
    return res


if __name__ == "__main__":
    for (B, N, H, rope) in [(2, 17, 2, True), (1, 197, 2, True), (2, 65, 1, True), (1, 200, 1, False), (2, 64, 1, False)]:
        r = check(B, N, H, rope)
        print(f"B={B} N={N} H={H} rope={rope}: " + "  ".join(f"{k} rel-L2 {v:.3e}" for k, v in r.items()))
        assert all(v < 2e-2 for v in r.values()), r
    print("dataflow of attention_bwd_tc.cu reproduces autograd")
