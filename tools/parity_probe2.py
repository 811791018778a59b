"""Stage-by-stage parity of ONE folded block: every kernel is fed the DEVICE's own inputs, the oracle recomputes the same
stage from those inputs (diagnostic for the 1e-3 parity budget)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import clipself_oracle as O, device_arith_oracle as DA
from clipself_b200.tower import TowerEngine, TowerCfg, stat_parts
from clipself_b200 import ops, _lib as L

which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
ocfg = O.CFG_TINY if which == "tiny" else O.CFG_B16
cfg = TowerCfg(image_size=ocfg.image_size, patch=ocfg.patch, width=ocfg.width, heads=ocfg.heads, layers=ocfg.layers,
               hidden=ocfg.hidden, embed_dim=ocfg.embed_dim, pt_seq_len=ocfg.pt_seq_len, ln_eps=ocfg.ln_eps)
dev = torch.device("cuda")
sd = O.synth_tower_weights(ocfg, 11)
images, boxes, crops = O.synth_batch(ocfg, 2, 4, 13, kind="grid")
imgs = crops.flatten(0, 1)
eng = TowerEngine(cfg, sd, dev)
n = imgs.shape[0]
ws = eng.workspace(n)
N, D = cfg.tokens, cfg.width
M = n * N
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
ar = DA.Arith(False)
bf = lambda t: t.to(torch.bfloat16).float()
cos, sin = O.rope_tables(cfg.grid, cfg.head_dim, cfg.pt_seq_len)
eps = cfg.ln_eps
with torch.no_grad():
    eng.embed(imgs.to(dev), ws.x, ws)
    for i in range(min(cfg.layers, 3)):
        pb = eng.w.blocks[i]
        p = f"blocks.{i}."
        x0 = ws.x[:M].cpu().clone().view(n, N, D)
        xb0 = ws.xb[:M].cpu().float().view(n, N, D)
        sx = (ws.stats_x, stat_parts(D), D, eps)
        print(f"block {i}: xb == bf16(x): {torch.equal(xb0, bf(x0))}")
        # --- qkv
        ops.gemm(ws.xb, pb.wqkv_f, ws.qkv, M=M, mode=L.EPI_QKV_ROPE, bias=pb.c2_qkv, rope=(eng.w.rope_pos, eng.w.rope_freq),
                 tokens=N, rope_cols=2 * D, ln_fold=(sx[0], pb.c1_qkv, *sx[1:]))
        Wqkv = torch.cat([sd[p + "attn.q_proj.weight"], sd[p + "attn.k_proj.weight"], sd[p + "attn.v_proj.weight"]])
        bqkv = torch.cat([sd[p + "attn.q_bias"], torch.zeros_like(sd[p + "attn.q_bias"]), sd[p + "attn.v_bias"]])
        pre = DA._folded_linear(ar, x0, xb0, Wqkv, sd[p + "norm1.weight"], sd[p + "norm1.bias"], bqkv, eps)
        H, hd = cfg.heads, 64
        q, k, v = pre.reshape(n, N, 3, H, hd).permute(2, 0, 3, 1, 4)
        q = torch.cat([q[:, :, :1], O.rope_apply(q[:, :, 1:], cos, sin)], dim=2)
        k = torch.cat([k[:, :, :1], O.rope_apply(k[:, :, 1:], cos, sin)], dim=2)
        qkv_ref = torch.stack([q, k, v]).permute(1, 3, 0, 2, 4).reshape(M, 3 * D)
        dq = ws.qkv[:M].cpu().float()
        print(f"   qkv  vs bf16(ref) {rel(dq, bf(qkv_ref)):.3e}   flips {(dq != bf(qkv_ref)).float().mean().item():.4f}  prernd rel {rel(dq, qkv_ref):.3e}")
        # --- attention from the DEVICE qkv
        ops.attention_fwd(ws.qkv, n, N, H, eng.scale, ws.att, row_stats=ws.stats_att)
        dqkv = dq.view(n, N, 3, H, hd).permute(2, 0, 3, 1, 4)
        s = dqkv[0] @ dqkv[1].transpose(-2, -1)
        sl2 = hd ** -0.5 * DA.LOG2E
        pp = torch.exp2(s * sl2 - s.amax(-1, keepdim=True) * sl2)
        o = ((bf(pp) @ dqkv[2]) / pp.sum(-1, keepdim=True)).transpose(1, 2).reshape(n, N, D)
        da = ws.att[:M].cpu().float().view(n, N, D)
        st = ws.stats_att[:M].cpu().view(n, N, 4 * H, 2).sum(2)
        print(f"   att  vs bf16(ref) {rel(da, bf(o)):.3e}   flips {(da != bf(o)).float().mean().item():.4f}   stats s1 {rel(st[..., 0], o.sum(-1)):.2e} s2 {rel(st[..., 1], (o * o).sum(-1)):.2e}")
        # --- proj from the DEVICE att + stats
        xin = ws.x[:M].clone()
        ops.gemm(ws.att, pb.wproj_f, ws.x, M=M, bias=pb.c2_proj, residual=ws.x, out2=ws.xb, stats_out=ws.stats_x,
                 ln_fold=(ws.stats_att, pb.c1_proj, 4 * H, D, eps))
        Wf = bf(sd[p + "attn.proj.weight"] * sd[p + "attn.inner_attn_ln.weight"][None, :])
        c1 = Wf.sum(1); c2 = sd[p + "attn.proj.weight"] @ sd[p + "attn.inner_attn_ln.bias"] + sd[p + "attn.proj.bias"]
        mean = st[..., 0:1] / D; var = (st[..., 1:2] / D - mean * mean).clamp_min(0); rstd = torch.rsqrt(var + eps)
        x1_ref = xin.cpu().view(n, N, D) + rstd * (da @ Wf.t()) - rstd * mean * c1 + c2
        x1 = ws.x[:M].cpu().view(n, N, D)
        delta_scale = (x1_ref - xin.cpu().view(n, N, D)).norm() / x1_ref.norm()
        print(f"   proj x rel {rel(x1, x1_ref):.3e}  (|delta|/|x| {delta_scale:.2f})  xb==bf16(x) {torch.equal(ws.xb[:M].cpu().float().view(n, N, D), bf(x1))}"
              f"  stats_x s2 {rel(ws.stats_x[:M].cpu().view(n, N, -1, 2).sum(2)[..., 1], (x1 * x1).sum(-1)):.2e}")
        # --- w12 from the device xb / stats
        ops.gemm(ws.xb, pb.w12_f, ws.h, M=M, mode=L.EPI_SWIGLU, bias=pb.c2_w12, stats_out=ws.stats_h,
                 ln_fold=(ws.stats_x, pb.c1_w12, stat_parts(D), D, eps))
        xb1 = ws.xb[:M].cpu().float().view(n, N, D)
        g2, b2 = sd[p + "norm2.weight"], sd[p + "norm2.bias"]
        gte = DA._folded_linear(ar, x1, xb1, sd[p + "mlp.w1.weight"], g2, b2, sd[p + "mlp.w1.bias"], eps)
        up = DA._folded_linear(ar, x1, xb1, sd[p + "mlp.w2.weight"], g2, b2, sd[p + "mlp.w2.bias"], eps)
        h_ref = gte / (1.0 + torch.exp(-gte)) * up
        dh = ws.h[:M, :cfg.hidden].cpu().float().view(n, N, -1)
        sh = ws.stats_h[:M].cpu().view(n, N, -1, 2).sum(2)
        print(f"   h    vs bf16(ref) {rel(dh, bf(h_ref)):.3e}   flips {(dh != bf(h_ref)).float().mean().item():.4f}  prernd rel {rel(dh, h_ref):.3e}  stats s2 {rel(sh[..., 1], (h_ref * h_ref).sum(-1)):.2e}")
        # --- w3 from the device h + stats
        xin = ws.x[:M].clone()
        ops.gemm(ws.h, pb.w3_f, ws.x, M=M, bias=pb.c2_w3, residual=ws.x, out2=ws.xb, stats_out=ws.stats_x,
                 ln_fold=(ws.stats_h, pb.c1_w3, cfg.hidden_pad // 64, cfg.hidden, eps))
        Wf = bf(sd[p + "mlp.w3.weight"] * sd[p + "mlp.ffn_ln.weight"][None, :])
        c1 = Wf.sum(1); c2 = sd[p + "mlp.w3.weight"] @ sd[p + "mlp.ffn_ln.bias"] + sd[p + "mlp.w3.bias"]
        Hd = cfg.hidden
        mean = sh[..., 0:1] / Hd; var = (sh[..., 1:2] / Hd - mean * mean).clamp_min(0); rstd = torch.rsqrt(var + eps)
        x2_ref = xin.cpu().view(n, N, D) + rstd * (dh @ Wf.t()) - rstd * mean * c1 + c2
        x2 = ws.x[:M].cpu().view(n, N, D)
        print(f"   w3   x rel {rel(x2, x2_ref):.3e}  (|delta|/|x| {((x2_ref - xin.cpu().view(n, N, D)).norm() / x2_ref.norm()).item():.2f})")
