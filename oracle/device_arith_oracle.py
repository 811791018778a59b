"""Device-arithmetic oracle.  TEST INFRASTRUCTURE ONLY (same rules as clipself_oracle.py).

`clipself_oracle.py` restates the reference in fp32 and is pinned by reference-generated fixtures.  This
module restates the SAME path with the rounding points of the B200 kernels made explicit, so
that the GPU parity tests can hold features to the 1e-3 that `north_star` names instead of to the size of
the bf16 noise itself (the reference's own bf16-autocast run deviates from its fp32 run by ~1.1e-2 rel-L2 on
the dense map):

  * every tensor-core operand is rounded to bf16 (round-to-nearest-even) exactly where the kernels round it
    — im2col patches, packed weights, LayerNorm outputs, q|k|v after RoPE, the softmax probabilities P, the
    attention output, the SwiGLU product, the bf16 copy of the residual stream;
  * accumulation, LayerNorm, softmax, RoPE, normalise, RoIAlign and the loss stay in f32;
  * `fold=True` is the frozen-tower pipeline (tower.py: block_inplace, folded form): LayerNorm folded into the
    following GEMM, y = rstd * (bf16(a) @ bf16(W diag(gamma))^T - mean * c1) + c2 with row statistics taken
    from the f32 values BEFORE they are rounded (raw moments E[a^2] - mean^2, as the epilogues do);
  * `fold=False` is the taped student pipeline (student.py): explicit LayerNorm kernels (two-pass centred
    variance) whose bf16 outputs feed the GEMMs, w1|w2 output stored as bf16 before the SiLU product.

Rounding is applied with a straight-through estimator, so autograd through this module gives the gradient of
the fp32 computation evaluated at the device's forward values: the remaining difference to the device's
gradients is the bf16 rounding of the BACKWARD operands only.

It is anchored on clipself_oracle.py: with rounding disabled (`exact=True`) every function here must
reproduce the fp32 oracle (tests/test_oracle_vs_golden.py), which in turn is pinned by the reference
fixtures.  Reference citations as in clipself_oracle.py (E = eva_vit_model.py, C = clipself.py).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import clipself_oracle as O

Tensor = torch.Tensor
LOG2E = 1.4426950408889634


class Arith:
    """Rounding policy: exact=True turns every rounding point into the identity (fp32 oracle)."""

    def __init__(self, exact: bool = False):
        self.exact = exact

    def r(self, x: Tensor) -> Tensor:
        """round to bf16, straight-through for autograd"""
        if self.exact:
            return x
        return x + (x.detach().to(torch.bfloat16).to(x.dtype) - x.detach())


def _moments_fold(a: Tensor, eps: float):
    """(rstd, rstd*mean) from raw moments of the f32 values, as the GEMM epilogues rebuild them from partial sums."""
    D = a.shape[-1]
    mean = a.sum(-1, keepdim=True) / D
    var = ((a * a).sum(-1, keepdim=True) / D - mean * mean).clamp_min(0.0)
    rstd = torch.rsqrt(var + eps)
    return rstd, rstd * mean


def _folded_linear(ar: Arith, a_stats: Tensor, a_op: Tensor, W: Tensor, gamma: Tensor, beta: Tensor, b, eps: float) -> Tensor:
    """LN(a) W^T + b in the folded form of cs_gemm_epilogue_t.ln_*: a_stats = the f32 values the statistics come
    from, a_op = their bf16 copy (the GEMM operand)."""
    Wf = ar.r(W * gamma[None, :])
    c1 = Wf.sum(1)
    c2 = W @ beta + (b if b is not None else 0.0)
    rstd, rm = _moments_fold(a_stats, eps)
    return rstd * (a_op @ Wf.t()) - rm * c1 + c2


def _ln_explicit(ar: Arith, x: Tensor, sd, name: str, eps: float) -> Tensor:
    """layernorm_fwd kernel: f32 statistics (two-pass), bf16 output."""
    w = sd[name + ".weight"]
    return ar.r(F.layer_norm(x, (w.numel(),), w, sd[name + ".bias"], eps))


def embed_tokens(ar: Arith, sd, images: Tensor, cfg) -> Tensor:
    """im2col (bf16) x bf16 conv weight, f32 accumulate, + bias + pos_embed in f32 (E:350-356, 540-544)."""
    x = F.conv2d(ar.r(images), ar.r(sd["patch_embed.proj.weight"]), sd["patch_embed.proj.bias"], stride=cfg.patch)
    x = x.flatten(2).transpose(1, 2)
    pos = O.rescale_pos_embed(sd["pos_embed"], images.shape[-1] // cfg.patch)
    x = x + pos[:, 1:]
    cls = (sd["cls_token"] + pos[:, :1]).expand(x.shape[0], -1, -1)
    return torch.cat([cls, x], dim=1)


def _attention_core(ar: Arith, qkv: Tensor, cfg, cos: Tensor, sin: Tensor):
    """qkv [B,N,3D] f32 (projection outputs incl. bias) -> RoPE (f32) -> bf16 -> softmax(q k^T / 8) v.
    Returns (att_f32 [B,N,D] before its bf16 rounding, for the row statistics)."""
    B, N, _ = qkv.shape
    H, hd = cfg.heads, cfg.head_dim
    q, k, v = qkv.reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    q = torch.cat([q[:, :, :1], O.rope_apply(q[:, :, 1:], cos, sin)], dim=2)
    k = torch.cat([k[:, :, :1], O.rope_apply(k[:, :, 1:], cos, sin)], dim=2)
    o = softmax_pv(ar, ar.r(q), ar.r(k), ar.r(v))
    return o.transpose(1, 2).reshape(B, N, H * hd)


def softmax_pv(ar: Arith, q: Tensor, k: Tensor, v: Tensor, block: int = 128) -> Tensor:
    """softmax(q k^T / sqrt(hd)) v on bf16 q, k, v [..., N, hd] with the kernels' rounding points.
    N <= 208 (attention_tc2.cu): one pass, P = ex2(s*c - max*c) rounded to bf16 for the MMA, and the row sum is taken by the
    tensor core from the SAME rounded P (P times a block of ones), i.e. the exact normaliser of the P V product.
    N  > 208 (attention_tc_long.cu): keys stream in blocks of 128 with the online softmax: P of a block is taken relative
    to the RUNNING maximum (so its bf16 rounding differs from the one-pass form), the output and the row sum are rescaled
    by 2^((m_old - m_new) c) in f32."""
    N, hd = q.shape[-2], q.shape[-1]
    c = hd ** -0.5 * LOG2E
    s = q @ k.transpose(-2, -1)                                  # f32 accumulate of bf16 products
    if N <= 208:
        p = ar.r(torch.exp2(s * c - s.amax(-1, keepdim=True) * c))
        return (p @ v) / p.sum(-1, keepdim=True)
    m_run = torch.full(s.shape[:-1] + (1,), float("-inf"))
    l_run = torch.zeros_like(m_run)
    o_run = torch.zeros(q.shape[:-1] + (v.shape[-1],))
    for j in range(0, N, block):
        sb = s[..., j:j + block]
        m_new = torch.maximum(m_run, sb.amax(-1, keepdim=True))
        p = torch.exp2(sb * c - m_new * c)
        beta = torch.exp2((m_run - m_new) * c)                   # 0 on the first block (m_run = -inf)
        o_run = o_run * beta + ar.r(p) @ v[..., j:j + block, :]
        l_run = l_run * beta + p.sum(-1, keepdim=True)
        m_run = m_new
    return o_run / l_run


def block_fold(ar: Arith, x: Tensor, sd, i: int, cfg, cos, sin, with_attention: bool = True) -> Tensor:
    """tower.py: block_inplace, fully folded form (5 launches per block)."""
    p = f"blocks.{i}."
    eps = cfg.ln_eps
    D = cfg.width
    xb = ar.r(x)
    Wqkv = torch.cat([sd[p + "attn.q_proj.weight"], sd[p + "attn.k_proj.weight"], sd[p + "attn.v_proj.weight"]])
    bqkv = torch.cat([sd[p + "attn.q_bias"], torch.zeros_like(sd[p + "attn.q_bias"]), sd[p + "attn.v_bias"]])
    g1, b1 = sd[p + "norm1.weight"], sd[p + "norm1.bias"]
    if with_attention:
        qkv = _folded_linear(ar, x, xb, Wqkv, g1, b1, bqkv, eps)
        att = _attention_core(ar, qkv, cfg, cos, sin)
        delta = _folded_linear(ar, att, ar.r(att), sd[p + "attn.proj.weight"], sd[p + "attn.inner_attn_ln.weight"],
                               sd[p + "attn.inner_attn_ln.bias"], sd[p + "attn.proj.bias"], eps)
    else:       # value-only block (E:317-324, 249-256): folded v-projection, explicit inner LN
        a = ar.r(_folded_linear(ar, x, xb, Wqkv[2 * D:], g1, b1, bqkv[2 * D:], eps))
        u = _ln_explicit(ar, a, sd, p + "attn.inner_attn_ln", eps)
        delta = u @ ar.r(sd[p + "attn.proj.weight"]).t() + sd[p + "attn.proj.bias"]
    x = x + delta
    xb = ar.r(x)
    g2, b2 = sd[p + "norm2.weight"], sd[p + "norm2.bias"]
    gte = _folded_linear(ar, x, xb, sd[p + "mlp.w1.weight"], g2, b2, sd[p + "mlp.w1.bias"], eps)
    up = _folded_linear(ar, x, xb, sd[p + "mlp.w2.weight"], g2, b2, sd[p + "mlp.w2.bias"], eps)
    h = gte / (1.0 + torch.exp(-gte)) * up                       # SWIGLU epilogue, f32
    delta = _folded_linear(ar, h, ar.r(h), sd[p + "mlp.w3.weight"], sd[p + "mlp.ffn_ln.weight"], sd[p + "mlp.ffn_ln.bias"],
                           sd[p + "mlp.w3.bias"], eps)
    return x + delta


def block_explicit(ar: Arith, x: Tensor, sd, i: int, cfg, cos, sin, with_attention: bool = True) -> Tensor:
    """student.py: forward (explicit LayerNorm kernels, everything kept on the tape)."""
    p = f"blocks.{i}."
    eps = cfg.ln_eps
    u = _ln_explicit(ar, x, sd, p + "norm1", eps)
    if with_attention:
        Wqkv = torch.cat([sd[p + "attn.q_proj.weight"], sd[p + "attn.k_proj.weight"], sd[p + "attn.v_proj.weight"]])
        bqkv = torch.cat([sd[p + "attn.q_bias"], torch.zeros_like(sd[p + "attn.q_bias"]), sd[p + "attn.v_bias"]])
        att = ar.r(_attention_core(ar, u @ ar.r(Wqkv).t() + bqkv, cfg, cos, sin))
    else:
        att = ar.r(u @ ar.r(sd[p + "attn.v_proj.weight"]).t() + sd[p + "attn.v_bias"])
    aln = _ln_explicit(ar, att, sd, p + "attn.inner_attn_ln", eps)
    xmid = x + aln @ ar.r(sd[p + "attn.proj.weight"]).t() + sd[p + "attn.proj.bias"]
    u2 = _ln_explicit(ar, xmid, sd, p + "norm2", eps)
    x1 = ar.r(u2 @ ar.r(sd[p + "mlp.w1.weight"]).t() + sd[p + "mlp.w1.bias"])      # w1|w2 output stored as bf16
    x2 = ar.r(u2 @ ar.r(sd[p + "mlp.w2.weight"]).t() + sd[p + "mlp.w2.bias"])
    h = ar.r(x1 / (1.0 + torch.exp(-x1)) * x2)                                      # swiglu_fwd kernel
    hln = _ln_explicit(ar, h, sd, p + "mlp.ffn_ln", eps)
    return xmid + hln @ ar.r(sd[p + "mlp.w3.weight"]).t() + sd[p + "mlp.w3.bias"]


def _block(ar, fold):
    return block_fold if fold else block_explicit


def tower_forward_cls(sd: Dict[str, Tensor], images: Tensor, cfg, fold: bool = True, exact: bool = False) -> Tensor:
    """Teacher path (E:533-586) with the device's rounding points."""
    ar = Arith(exact)
    cos, sin = O.rope_tables(O.image_grid(images, cfg), cfg.head_dim, cfg.pt_seq_len)
    x = embed_tokens(ar, sd, images, cfg)
    for i in range(cfg.layers):
        x = _block(ar, fold)(ar, x, sd, i, cfg, cos, sin)
    u = _ln_explicit(ar, x[:, 0], sd, "norm", cfg.ln_eps)
    return u @ ar.r(sd["head.weight"]).t() + sd["head.bias"]


def tower_encode_dense(sd: Dict[str, Tensor], images: Tensor, cfg, fold: bool = False, exact: bool = False) -> Tensor:
    """Student dense map, NHWC, unit-norm per token (E:588-623) with the device's rounding points."""
    ar = Arith(exact)
    g = O.image_grid(images, cfg)
    cos, sin = O.rope_tables(g, cfg.head_dim, cfg.pt_seq_len)
    x = embed_tokens(ar, sd, images, cfg)
    for i in range(cfg.layers - 1):
        x = _block(ar, fold)(ar, x, sd, i, cfg, cos, sin)
    x = _block(ar, fold)(ar, x, sd, cfg.layers - 1, cfg, cos, sin, with_attention=False)[:, 1:]
    u = _ln_explicit(ar, x, sd, "norm", cfg.ln_eps)
    y = u @ ar.r(sd["head.weight"]).t() + sd["head.bias"]
    return F.normalize(y, dim=-1).reshape(images.shape[0], g, g, -1)


def clipself_step(student_sd, teacher_sd, images, normed_boxes, image_crops, cfg, cosine_weight: float = 1.0,
                  exact: bool = False):
    """C:7-49 with the device's arithmetic: folded frozen teacher, taped student, f32 region path."""
    rois, crop_idx = O.extract_rois(normed_boxes)
    crops = image_crops.flatten(0, 1)[crop_idx]
    with torch.no_grad():
        teacher = tower_forward_cls(teacher_sd, crops, cfg, fold=True, exact=exact)
    dense = tower_encode_dense(student_sd, images, cfg, fold=False, exact=exact)
    boxes = O.denormalize_boxes(rois, dense.shape[1], dense.shape[2])
    student = O.roi_align_1x1_nhwc(dense, boxes)
    loss = O.cosine_loss(student, teacher, cosine_weight)
    return dict(loss=loss, student_roi=student, teacher=teacher, dense=dense)
