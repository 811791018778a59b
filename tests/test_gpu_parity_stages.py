"""Stage-wise parity at north_star's tolerance (1e-3 rel) against the device-arithmetic oracle.

Why stage-wise: two implementations of a bf16 pipeline that are not bit-identical cannot agree to 1e-3 END TO END.
A difference e in a value that is about to be rounded to bf16 flips the rounding of a fraction ~e/ulp of the elements,
each flip is worth one ulp (2^-8 relative), so the rounded tensors differ by about sqrt(e * ulp): the fixed point of
e -> sqrt(e * ulp) is e ~ ulp/3 ~ 1.3e-3 per rounding point, whatever the initial difference was (f32 summation order
is enough), and the blocks then compound it (measured: 3e-3 .. 6e-3 on the 12-block dense map, the reference's own
bf16-autocast run is at 1.1e-2 from its fp32 run).  So the kernels are held to the tolerance where it is attainable
and meaningful: every stage is fed the DEVICE's own inputs, the oracle (oracle/device_arith_oracle.py, anchored on the
fp32 oracle and through it on the reference fixtures) recomputes that stage from the same inputs, and the outputs must
agree to 2.5e-4 (bf16 outputs; measured <= 1e-4, i.e. a few rounding flips) resp. 1e-5 (f32 outputs).  A kernel whose
arithmetic is off by 0.1 % fails here; tests/test_gpu_tower.py and test_gpu_step.py keep the end-to-end bounds.
"""
import pytest
import torch

from oracle import clipself_oracle as O
from oracle import device_arith_oracle as DA

pytestmark = pytest.mark.gpu

TOL_BF16, TOL_F32 = 2.5e-4, 1e-5


def _cfg(o):
    from clipself_b200.tower import TowerCfg
    return TowerCfg(image_size=o.image_size, patch=o.patch, width=o.width, heads=o.heads, layers=o.layers,
                    hidden=o.hidden, embed_dim=o.embed_dim, pt_seq_len=o.pt_seq_len, ln_eps=o.ln_eps)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def bf(t):
    return t.to(torch.bfloat16).float()


def _attention_ref(qkv, n, N, H, hd=64):
    """softmax(q k^T / 8) v from bf16 q|k|v with the kernels' rounding points (P rounded for the MMA, f32 row sum)."""
    q, k, v = qkv.view(n, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    return DA.softmax_pv(DA.Arith(False), q, k, v).transpose(1, 2).reshape(n, N, H * hd)


def _moments(stats, dim):
    mean = stats[..., 0:1] / dim
    var = (stats[..., 1:2] / dim - mean * mean).clamp_min(0)
    return mean, torch.rsqrt(var + 1e-6)


@pytest.mark.parametrize("which,blocks", [("tiny", (0, 1, 2)), ("b16", (0, 1, 11)), ("l14", (0, 23))])
def test_frozen_tower_stages(which, blocks):
    """Every launch of the LayerNorm-folded frozen-tower block (tower.py: block_inplace) on the device's own inputs."""
    from clipself_b200 import _lib as L, ops
    from clipself_b200.tower import TowerEngine, stat_parts
    ocfg = {"tiny": O.CFG_TINY, "b16": O.CFG_B16, "l14": O.CFG_L14_336}[which]
    cfg = _cfg(ocfg)
    dev = torch.device("cuda")
    sd = O.synth_tower_weights(ocfg, 11)
    _, _, crops = O.synth_batch(ocfg, 1 if which == "l14" else 2, 2 if which == "l14" else 4, 13, kind="grid")
    imgs = crops.flatten(0, 1)
    eng = TowerEngine(cfg, sd, dev)
    n, N, D, H, Hd = imgs.shape[0], cfg.tokens, cfg.width, cfg.heads, cfg.hidden
    M = n * N
    ws = eng.workspace(n)
    ar = DA.Arith(False)
    cos, sin = O.rope_tables(cfg.grid, 64, cfg.pt_seq_len)
    eps = cfg.ln_eps
    worst = {}

    def check(name, got, ref, tol):
        r = rel(got, ref)
        worst[name] = max(worst.get(name, 0.0), r)
        assert r <= tol, (which, name, r)

    with torch.no_grad():
        eng.embed(imgs.to(dev), ws.x, ws)
        x_ref = DA.embed_tokens(ar, sd, imgs, ocfg)
        check("embed", ws.x[:M].cpu().view_as(x_ref), x_ref, TOL_F32)
        for i in range(cfg.layers):
            if i not in blocks:
                eng.block_inplace(i, ws, n)
                continue
            pb, p = eng.w.blocks[i], f"blocks.{i}."
            x0 = ws.x[:M].cpu().clone().view(n, N, D)
            xb0 = ws.xb[:M].cpu().float().view(n, N, D)
            assert torch.equal(xb0, bf(x0)), "the bf16 copy of the residual stream is its round-to-nearest-even"
            sx = ws.stats_x[:M].cpu().view(n, N, -1, 2).sum(2)
            check("stats_x", sx[..., 1], (x0 * x0).sum(-1), TOL_F32)
            fold = (ws.stats_x, stat_parts(D), D, eps)
            # ---- q|k|v with norm1 folded + RoPE
            ops.gemm(ws.xb, pb.wqkv_f, ws.qkv, M=M, mode=L.EPI_QKV_ROPE, bias=pb.c2_qkv, rope=(eng.w.rope_pos, eng.w.rope_freq),
                     tokens=N, rope_cols=2 * D, ln_fold=(fold[0], pb.c1_qkv, *fold[1:]))
            Wqkv = torch.cat([sd[p + "attn.q_proj.weight"], sd[p + "attn.k_proj.weight"], sd[p + "attn.v_proj.weight"]])
            bqkv = torch.cat([sd[p + "attn.q_bias"], torch.zeros_like(sd[p + "attn.q_bias"]), sd[p + "attn.v_bias"]])
            pre = DA._folded_linear(ar, x0, xb0, Wqkv, sd[p + "norm1.weight"], sd[p + "norm1.bias"], bqkv, eps)
            q, k, v = pre.reshape(n, N, 3, H, 64).permute(2, 0, 3, 1, 4)
            q = torch.cat([q[:, :, :1], O.rope_apply(q[:, :, 1:], cos, sin)], dim=2)
            k = torch.cat([k[:, :, :1], O.rope_apply(k[:, :, 1:], cos, sin)], dim=2)
            qkv_ref = torch.stack([q, k, v]).permute(1, 3, 0, 2, 4).reshape(M, 3 * D)
            dq = ws.qkv[:M].cpu().float()
            check("qkv", dq, bf(qkv_ref), TOL_BF16)
            # ---- attention
            ops.attention_fwd(ws.qkv, n, N, H, eng.scale, ws.att, row_stats=ws.stats_att)
            o = _attention_ref(dq, n, N, H)
            da = ws.att[:M].cpu().float().view(n, N, D)
            check("attention", da, bf(o), TOL_BF16)
            st = ws.stats_att[:M].cpu().view(n, N, 4 * H, 2).sum(2)
            check("stats_att", st[..., 1], (o * o).sum(-1), 2e-5)
            # ---- proj with inner_attn_ln folded, residual add, new stream emitted as f32 + bf16 + statistics
            xin = ws.x[:M].cpu().clone().view(n, N, D)
            ops.gemm(ws.att, pb.wproj_f, ws.x, M=M, bias=pb.c2_proj, residual=ws.x, out2=ws.xb, stats_out=ws.stats_x,
                     ln_fold=(ws.stats_att, pb.c1_proj, 4 * H, D, eps))
            Wf = bf(sd[p + "attn.proj.weight"] * sd[p + "attn.inner_attn_ln.weight"][None, :])
            c2 = sd[p + "attn.proj.weight"] @ sd[p + "attn.inner_attn_ln.bias"] + sd[p + "attn.proj.bias"]
            mean, rstd = _moments(st, D)
            x1_ref = xin + rstd * (da @ Wf.t()) - rstd * mean * Wf.sum(1) + c2
            x1 = ws.x[:M].cpu().clone().view(n, N, D)
            check("proj", x1, x1_ref, TOL_F32)
            xb1 = ws.xb[:M].cpu().float().view(n, N, D)
            assert torch.equal(xb1, bf(x1))
            # ---- w1|w2 with norm2 folded + SiLU*mul
            ops.gemm(ws.xb, pb.w12_f, ws.h, M=M, mode=L.EPI_SWIGLU, bias=pb.c2_w12, stats_out=ws.stats_h,
                     ln_fold=(ws.stats_x, pb.c1_w12, stat_parts(D), D, eps))
            g2, b2 = sd[p + "norm2.weight"], sd[p + "norm2.bias"]
            gte = DA._folded_linear(ar, x1, xb1, sd[p + "mlp.w1.weight"], g2, b2, sd[p + "mlp.w1.bias"], eps)
            up = DA._folded_linear(ar, x1, xb1, sd[p + "mlp.w2.weight"], g2, b2, sd[p + "mlp.w2.bias"], eps)
            h_ref = gte / (1.0 + torch.exp(-gte)) * up
            dh = ws.h[:M, :Hd].cpu().float().view(n, N, Hd)
            check("swiglu", dh, bf(h_ref), TOL_BF16)
            sh = ws.stats_h[:M].cpu().view(n, N, -1, 2).sum(2)
            check("stats_h", sh[..., 1], (h_ref * h_ref).sum(-1), 2e-5)
            # ---- w3 with ffn_ln folded
            ops.gemm(ws.h, pb.w3_f, ws.x, M=M, bias=pb.c2_w3, residual=ws.x, out2=ws.xb, stats_out=ws.stats_x,
                     ln_fold=(ws.stats_h, pb.c1_w3, cfg.hidden_pad // 64, Hd, eps))
            Wf = bf(sd[p + "mlp.w3.weight"] * sd[p + "mlp.ffn_ln.weight"][None, :])
            c2 = sd[p + "mlp.w3.weight"] @ sd[p + "mlp.ffn_ln.bias"] + sd[p + "mlp.w3.bias"]
            mean, rstd = _moments(sh, Hd)
            x2_ref = x1 + rstd * (dh @ Wf.t()) - rstd * mean * Wf.sum(1) + c2
            check("w3", ws.x[:M].cpu().view(n, N, D), x2_ref, TOL_F32)
    print(f"{which}: worst stage rel-L2 " + ", ".join(f"{k} {v:.1e}" for k, v in worst.items()))


@pytest.mark.parametrize("which", ["tiny", "b16"])
def test_student_forward_stages_from_the_tape(which):
    """The taped student forward (student.py: explicit LayerNorm kernels): every saved activation against the oracle
    stage computed from the tape's own inputs of that stage."""
    from clipself_b200.student import StudentEngine
    ocfg = {"tiny": O.CFG_TINY, "b16": O.CFG_B16}[which]
    cfg = _cfg(ocfg)
    dev = torch.device("cuda")
    sd = O.synth_tower_weights(ocfg, 21)
    images, _, _ = O.synth_batch(ocfg, 2, 4, 23, kind="grid")
    eng = StudentEngine(cfg, {k: v.to(dev) for k, v in sd.items()}, dev)
    with torch.no_grad():
        dense = eng.forward(images.to(dev))
    t = eng._tape
    n, N, D, H, Hd, Lr = images.shape[0], cfg.tokens, cfg.width, cfg.heads, cfg.hidden, cfg.layers
    ar = DA.Arith(False)
    cos, sin = O.rope_tables(cfg.grid, 64, cfg.pt_seq_len)
    eps = cfg.ln_eps
    worst = {}

    def check(name, got, ref, tol):
        r = rel(got, ref)
        worst[name] = max(worst.get(name, 0.0), r)
        assert r <= tol, (which, name, r)

    c = lambda z: z.cpu().float()  # noqa: E731
    with torch.no_grad():
        check("embed", c(t.x[0]).view(n, N, D), DA.embed_tokens(ar, sd, images, ocfg), TOL_F32)
        for i in (range(Lr) if which == "tiny" else (0, 5, Lr - 2, Lr - 1)):
            p = f"blocks.{i}."
            x, u = c(t.x[i]), c(t.u[i])
            check("norm1", u, DA._ln_explicit(ar, x, sd, p + "norm1", eps), TOL_BF16)
            if i < Lr - 1:
                Wqkv = torch.cat([sd[p + "attn.q_proj.weight"], sd[p + "attn.k_proj.weight"], sd[p + "attn.v_proj.weight"]])
                bqkv = torch.cat([sd[p + "attn.q_bias"], torch.zeros_like(sd[p + "attn.q_bias"]), sd[p + "attn.v_bias"]])
                pre = (u @ bf(Wqkv).t() + bqkv).view(n, N, 3 * D)
                q, k, v = pre.reshape(n, N, 3, H, 64).permute(2, 0, 3, 1, 4)
                q = torch.cat([q[:, :, :1], O.rope_apply(q[:, :, 1:], cos, sin)], dim=2)
                k = torch.cat([k[:, :, :1], O.rope_apply(k[:, :, 1:], cos, sin)], dim=2)
                qkv_ref = torch.stack([q, k, v]).permute(1, 3, 0, 2, 4).reshape(n * N, 3 * D)
                dq = c(t.qkv[i])
                check("qkv", dq, bf(qkv_ref), TOL_BF16)
                check("attention", c(t.att[i]).view(n, N, D), bf(_attention_ref(dq, n, N, H)), TOL_BF16)
            else:
                check("v_only", c(t.att[i]), bf(u @ bf(sd[p + "attn.v_proj.weight"]).t() + sd[p + "attn.v_bias"]), TOL_BF16)
            att, aln = c(t.att[i]), c(t.aln[i])
            check("inner_ln", aln, DA._ln_explicit(ar, att, sd, p + "attn.inner_attn_ln", eps), TOL_BF16)
            xmid = c(t.xmid[i])
            check("proj", xmid, x + aln @ bf(sd[p + "attn.proj.weight"]).t() + sd[p + "attn.proj.bias"], TOL_F32)
            u2 = c(t.u2[i])
            check("norm2", u2, DA._ln_explicit(ar, xmid, sd, p + "norm2", eps), TOL_BF16)
            Hp = cfg.hidden_pad
            x12 = c(t.x12[i])
            x1, x2 = x12[:, :Hd], x12[:, Hp:Hp + Hd]
            check("w1", x1, bf(u2 @ bf(sd[p + "mlp.w1.weight"]).t() + sd[p + "mlp.w1.bias"]), TOL_BF16)
            check("w2", x2, bf(u2 @ bf(sd[p + "mlp.w2.weight"]).t() + sd[p + "mlp.w2.bias"]), TOL_BF16)
            h = c(t.h[i])[:, :Hd]
            check("swiglu", h, bf(x1 / (1.0 + torch.exp(-x1)) * x2), TOL_BF16)
            hln = c(t.hln[i])[:, :Hd]
            check("ffn_ln", hln, DA._ln_explicit(ar, h, sd, p + "mlp.ffn_ln", eps), TOL_BF16)
            check("w3", c(t.x[i + 1]), xmid + hln @ bf(sd[p + "mlp.w3.weight"]).t() + sd[p + "mlp.w3.bias"], TOL_F32)
        xl = c(t.x[Lr]).view(n, N, D)[:, 1:].reshape(-1, D)
        tok = c(t.tok_ln)
        check("final_ln", tok, DA._ln_explicit(ar, xl, sd, "norm", eps), TOL_BF16)
        head = tok @ bf(sd["head.weight"]).t() + sd["head.bias"]
        check("head", c(t.head), head, TOL_F32)
        check("dense", dense.cpu().reshape(-1, cfg.embed_dim), torch.nn.functional.normalize(head, dim=-1), TOL_F32)
    print(f"{which}: worst stage rel-L2 " + ", ".join(f"{k} {v:.1e}" for k, v in worst.items()))
