"""CPU oracle for the CLIPSelf distillation hot path.  TEST INFRASTRUCTURE ONLY.

This is a plain torch-fp32 (CPU) restatement of the arithmetic of the reference's hot path
(wusize/CLIPSelf @ 1c7fe9c).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package
``clipself_b200`` never does and fails loudly when its CUDA library is missing.

Pinning: the reference ships no tests / golden vectors, so this restatement is
pinned against outputs of the reference ITSELF, generated in the build container by
``tests/golden/make_golden.py`` (imports /root/reference/src with the shims in
``oracle/ref_stubs``) and committed as ``tests/golden/*.npz``;
``tests/test_oracle_vs_golden.py`` checks every tensor.

Every function cites the reference lines it restates (paths relative to the reference repo):
  E  = src/open_clip/eva_clip/eva_vit_model.py      R = src/open_clip/eva_clip/rope.py
  M  = src/open_clip/eva_clip/model.py              C = src/training/clipself.py
torchvision.ops.roi_align (unpinned dependency) is restated from its documented
semantics for the single configuration the path uses: output (1,1), spatial_scale 1.0,
sampling_ratio -1, aligned=True.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass(frozen=True)
class TowerCfg:
    """Architecture numbers of one EVA02 vision tower (E:399-405, model_configs/*.json)."""
    image_size: int = 224
    patch: int = 16
    width: int = 768
    heads: int = 12
    layers: int = 12
    hidden: int = 2048          # int(width * mlp_ratio)  (E:271)
    embed_dim: int = 512        # head out features
    pt_seq_len: int = 16        # rope pretrain grid (E:404, json pt_hw_seq_len)
    ln_eps: float = 1e-6        # M:123

    @property
    def grid(self) -> int:
        return self.image_size // self.patch

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + 1

    @property
    def head_dim(self) -> int:
        return self.width // self.heads


CFG_B16 = TowerCfg()
CFG_L14_336 = TowerCfg(image_size=336, patch=14, width=1024, heads=16, layers=24,
                       hidden=2730, embed_dim=768)
CFG_TINY = TowerCfg(image_size=64, patch=16, width=128, heads=2, layers=3, hidden=384,
                    embed_dim=64)


# --------------------------------------------------------------------------------------
# RoPE  (R:118-142 table construction; R:25-29 rotate_half; R:148-164 apply)
# --------------------------------------------------------------------------------------
def rope_tables(grid: int, head_dim: int, pt_seq_len: int = 16, theta: float = 10000.0
                ) -> Tuple[Tensor, Tensor]:
    """cos/sin tables [grid*grid, head_dim].

    R:118 freqs = 1/theta^(arange(0,dim,2)/dim) with dim = head_dim/2 (E:428-436);
    R:127 t = arange(ft)/ft*pt; R:129-131 outer product, each freq repeated twice
    (interleaved), rows and columns concatenated: first half indexed by the token's row,
    second half by its column.
    """
    dim = head_dim // 2
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    t = torch.arange(grid) / grid * pt_seq_len
    ang = t[:, None] * freqs[None, :]                      # [grid, dim/2]
    ang = ang.repeat_interleave(2, dim=-1)                 # [grid, dim]
    full = torch.cat([ang[:, None, :].expand(grid, grid, dim),
                      ang[None, :, :].expand(grid, grid, dim)], dim=-1)
    full = full.reshape(grid * grid, 2 * dim)
    return full.cos(), full.sin()


def rope_apply(t: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """t*cos + rot(t)*sin with rot: (x0,x1)->(-x1,x0) on interleaved pairs (R:25-29,164)."""
    pairs = t.reshape(*t.shape[:-1], -1, 2)
    rot = torch.stack((-pairs[..., 1], pairs[..., 0]), dim=-1).reshape(t.shape)
    return t * cos + rot * sin


# --------------------------------------------------------------------------------------
# Tower pieces
# --------------------------------------------------------------------------------------
def _ln(x: Tensor, sd: Dict[str, Tensor], name: str, eps: float) -> Tensor:
    w = sd[name + ".weight"]
    return F.layer_norm(x, (w.numel(),), w, sd[name + ".bias"], eps)


def embed_tokens(sd: Dict[str, Tensor], images: Tensor, cfg: TowerCfg) -> Tensor:
    """conv(p, stride p)+bias -> [B,hw,D]; prepend cls; add pos_embed (E:350-356, E:540-544)."""
    x = F.conv2d(images, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"],
                 stride=cfg.patch)
    x = x.flatten(2).transpose(1, 2)
    cls = sd["cls_token"].expand(x.shape[0], -1, -1)
    x = torch.cat([cls, x], dim=1)
    return x + rescale_pos_embed(sd["pos_embed"], images.shape[-1] // cfg.patch)


def rescale_pos_embed(pos_embed: Tensor, grid: int) -> Tensor:
    """E:631-643 (rescale_positional_embedding): unchanged at the pretraining grid; otherwise the CLS
    row is kept and the [g0,g0] grid of patch rows is resampled bicubically (align_corners=False)."""
    g0 = int(round((pos_embed.shape[1] - 1) ** 0.5))
    if g0 == grid:
        return pos_embed
    pe = pos_embed[0, 1:].T.contiguous().view(1, -1, g0, g0)
    pe = F.interpolate(pe, (grid, grid), mode="bicubic", align_corners=False).view(-1, grid * grid)
    return torch.cat([pos_embed[0, :1], pe.T.contiguous()], dim=0)[None]


def image_grid(images: Tensor, cfg: TowerCfg) -> int:
    """Token grid of a (square) input: the reference's towers accept any resolution — RoPE tables are
    regenerated with ft_seq_len = grid (rope.py:179-214) and pos_embed is rescaled (E:631-643)."""
    assert images.shape[-1] == images.shape[-2] and images.shape[-1] % cfg.patch == 0
    return images.shape[-1] // cfg.patch


def swiglu(x: Tensor, sd: Dict[str, Tensor], p: str, eps: float) -> Tensor:
    """w3(ffn_ln(silu(w1 x) * (w2 x))) (E:98-105)."""
    x1 = F.linear(x, sd[p + "w1.weight"], sd[p + "w1.bias"])
    x2 = F.linear(x, sd[p + "w2.weight"], sd[p + "w2.bias"])
    h = F.silu(x1) * x2
    h = _ln(h, sd, p + "ffn_ln", eps)
    return F.linear(h, sd[p + "w3.weight"], sd[p + "w3.bias"])


def attention(x: Tensor, sd: Dict[str, Tensor], p: str, cfg: TowerCfg,
              cos: Tensor, sin: Tensor) -> Tensor:
    """Sub-LN attention with 2-D RoPE on the patch tokens (E:174-247, math branch E:221-246)."""
    B, N, D = x.shape
    H, hd = cfg.heads, cfg.head_dim
    q = F.linear(x, sd[p + "q_proj.weight"], sd[p + "q_bias"])
    k = F.linear(x, sd[p + "k_proj.weight"], None)
    v = F.linear(x, sd[p + "v_proj.weight"], sd[p + "v_bias"])
    q = q.reshape(B, N, H, hd).permute(0, 2, 1, 3)
    k = k.reshape(B, N, H, hd).permute(0, 2, 1, 3)
    v = v.reshape(B, N, H, hd).permute(0, 2, 1, 3)
    q = torch.cat([q[:, :, :1], rope_apply(q[:, :, 1:], cos, sin)], dim=2)   # E:194-204
    k = torch.cat([k[:, :, :1], rope_apply(k[:, :, 1:], cos, sin)], dim=2)
    att = (q * hd ** -0.5) @ k.transpose(-2, -1)
    att = att.softmax(dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, N, D)
    o = _ln(o, sd, p + "inner_attn_ln", cfg.ln_eps)
    return F.linear(o, sd[p + "proj.weight"], sd[p + "proj.bias"])


def attention_value_only(x: Tensor, sd: Dict[str, Tensor], p: str, cfg: TowerCfg) -> Tensor:
    """v-proj -> inner LN -> out-proj, no token mixing (E:249-256)."""
    a = F.linear(x, sd[p + "v_proj.weight"], sd[p + "v_bias"])
    a = _ln(a, sd, p + "inner_attn_ln", cfg.ln_eps)
    return F.linear(a, sd[p + "proj.weight"], sd[p + "proj.bias"])


def block(x: Tensor, sd: Dict[str, Tensor], i: int, cfg: TowerCfg, cos: Tensor, sin: Tensor,
          with_attention: bool = True) -> Tensor:
    """Pre-norm residual block (E:300-307); value-only variant (E:317-324)."""
    p = f"blocks.{i}."
    u = _ln(x, sd, p + "norm1", cfg.ln_eps)
    if with_attention:
        x = x + attention(u, sd, p + "attn.", cfg, cos, sin)
    else:
        x = x + attention_value_only(u, sd, p + "attn.", cfg)
    return x + swiglu(_ln(x, sd, p + "norm2", cfg.ln_eps), sd, p + "mlp.", cfg.ln_eps)


def tower_forward_cls(sd: Dict[str, Tensor], images: Tensor, cfg: TowerCfg,
                      taps: Dict[str, Tensor] | None = None) -> Tensor:
    """Teacher path: all blocks, final LN, CLS row, head (E:533-570, E:581-586)."""
    cos, sin = rope_tables(image_grid(images, cfg), cfg.head_dim, cfg.pt_seq_len)
    x = embed_tokens(sd, images, cfg)
    if taps is not None:
        taps["tokens0"] = x
    for i in range(cfg.layers):
        x = block(x, sd, i, cfg, cos, sin)
        if taps is not None:
            taps[f"block{i}"] = x
    x = _ln(x, sd, "norm", cfg.ln_eps)[:, 0]
    return F.linear(x, sd["head.weight"], sd["head.bias"])


def block_cls_only(x: Tensor, sd: Dict[str, Tensor], i: int, cfg: TowerCfg, cos: Tensor, sin: Tensor) -> Tensor:
    """The CLS row of block(x, ...) computed WITHOUT the other rows' outputs: k and v of every token, one query row through
    attention, inner LN, proj, norm2 and the SwiGLU MLP (what csrc/tower.cu: block_cls_tail runs for the teacher's last block,
    since E:565-569 returns norm(x)[:, 0] only).  Returns [B, 1, D]; must equal block(...)[:, :1] (tests/test_oracle_vs_golden)."""
    p = f"blocks.{i}."
    B, N, D = x.shape
    H, hd = cfg.heads, cfg.head_dim
    u = _ln(x, sd, p + "norm1", cfg.ln_eps)
    a = p + "attn."
    q = F.linear(u[:, :1], sd[a + "q_proj.weight"], sd[a + "q_bias"]).reshape(B, 1, H, hd).permute(0, 2, 1, 3)   # CLS: no rotation (E:194-204)
    k = F.linear(u, sd[a + "k_proj.weight"], None).reshape(B, N, H, hd).permute(0, 2, 1, 3)
    v = F.linear(u, sd[a + "v_proj.weight"], sd[a + "v_bias"]).reshape(B, N, H, hd).permute(0, 2, 1, 3)
    k = torch.cat([k[:, :, :1], rope_apply(k[:, :, 1:], cos, sin)], dim=2)
    att = ((q * hd ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, 1, D)
    o = _ln(o, sd, a + "inner_attn_ln", cfg.ln_eps)
    xc = x[:, :1] + F.linear(o, sd[a + "proj.weight"], sd[a + "proj.bias"])
    return xc + swiglu(_ln(xc, sd, p + "norm2", cfg.ln_eps), sd, p + "mlp.", cfg.ln_eps)


def tower_forward_cls_tail(sd: Dict[str, Tensor], images: Tensor, cfg: TowerCfg) -> Tensor:
    """tower_forward_cls with the last block restricted to the CLS row (the form the device runs)."""
    cos, sin = rope_tables(image_grid(images, cfg), cfg.head_dim, cfg.pt_seq_len)
    x = embed_tokens(sd, images, cfg)
    for i in range(cfg.layers - 1):
        x = block(x, sd, i, cfg, cos, sin)
    x = block_cls_only(x, sd, cfg.layers - 1, cfg, cos, sin)
    x = _ln(x, sd, "norm", cfg.ln_eps)[:, 0]
    return F.linear(x, sd["head.weight"], sd["head.bias"])


def tower_encode_dense(sd: Dict[str, Tensor], images: Tensor, cfg: TowerCfg,
                       taps: Dict[str, Tensor] | None = None) -> Tensor:
    """Student dense map, NHWC [B,h,w,C], unit-norm per token (E:588-623)."""
    g = image_grid(images, cfg)
    cos, sin = rope_tables(g, cfg.head_dim, cfg.pt_seq_len)
    x = embed_tokens(sd, images, cfg)
    for i in range(cfg.layers - 1):
        x = block(x, sd, i, cfg, cos, sin)
        if taps is not None:
            taps[f"block{i}"] = x
    x = block(x, sd, cfg.layers - 1, cfg, cos, sin, with_attention=False)[:, 1:]
    x = _ln(x, sd, "norm", cfg.ln_eps)
    x = F.linear(x, sd["head.weight"], sd["head.bias"])
    x = F.normalize(x, dim=-1)
    return x.reshape(images.shape[0], g, g, -1)


# --------------------------------------------------------------------------------------
# Region path
# --------------------------------------------------------------------------------------
def extract_rois(normed_boxes: Tensor) -> Tuple[List[Tensor], Tensor]:
    """valid = boxes[..., 4] > 0.5; image-major gather (C:29-36).

    Returns (per-image [k_i,4] box lists, flat int64 index into the [B*K] crop axis)."""
    B, K, _ = normed_boxes.shape
    rois, idx = [], []
    for b in range(B):
        valid = normed_boxes[b, :, -1] > 0.5
        rois.append(normed_boxes[b, valid, :4])
        idx.append(torch.nonzero(valid).flatten() + b * K)
    return rois, torch.cat(idx) if idx else torch.zeros(0, dtype=torch.long)


def denormalize_boxes(rois: Sequence[Tensor], h: int, w: int) -> List[Tensor]:
    """x*=w, y*=h in the box dtype (E:655-664)."""
    out = []
    for r in rois:
        r = r.clone()
        r[:, [0, 2]] *= w
        r[:, [1, 3]] *= h
        out.append(r)
    return out


def roi_align_1x1_nhwc(fmap: Tensor, boxes: Sequence[Tensor]) -> Tensor:
    """RoIAlign with output 1x1, scale 1, adaptive sampling grid, aligned=True, on an NHWC map.

    fmap [B,H,W,C]; boxes: per-image [k,4] (x0,y0,x1,y1) in feature-map units.
    Semantics (torchvision roi_align, call site E:628-629): start = x0-0.5; roi_w = x1-x0
    (no clamp when aligned); grid = ceil(roi_w) x ceil(roi_h); sample (i+.5)*roi/grid;
    bilinear with: outside [-1,size] -> 0, clamp below at 0, last pixel clamps; mean over
    max(grid_h*grid_w,1).  Differentiable (pure torch indexing)."""
    B, H, W, C = fmap.shape
    outs = []
    for b, bx in enumerate(boxes):
        for r in range(bx.shape[0]):
            x0, y0, x1, y1 = (float(v) for v in bx[r])
            # fp32 arithmetic like the CUDA/CPU kernels (T = float)
            f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))  # noqa: E731
            sw, sh = f32(x0 - 0.5), f32(y0 - 0.5)
            rw, rh = f32(f32(x1 - 0.5) - sw), f32(f32(y1 - 0.5) - sh)
            gw, gh = int(math.ceil(rw)), int(math.ceil(rh))
            cnt = max(gw * gh, 1)
            acc = fmap.new_zeros(C)
            # torchvision: start + ph*bin + (i + .5f) * bin / grid  (ph = 0, bin = roi size)
            for iy in range(gh):
                y = f32(sh + f32(f32((iy + 0.5) * rh) / gh))
                for ix in range(gw):
                    x = f32(sw + f32(f32((ix + 0.5) * rw) / gw))
                    acc = acc + _bilinear(fmap[b], y, x, H, W)
            outs.append(acc / cnt)
    if not outs:
        return fmap.new_zeros(0, C)
    return torch.stack(outs)


def _bilinear(f: Tensor, y: float, x: float, H: int, W: int) -> Tensor:
    if y < -1.0 or y > H or x < -1.0 or x > W:
        return f.new_zeros(f.shape[-1])
    y = max(y, 0.0)
    x = max(x, 0.0)
    yl, xl = int(y), int(x)
    if yl >= H - 1:
        yh = yl = H - 1
        y = float(yl)
    else:
        yh = yl + 1
    if xl >= W - 1:
        xh = xl = W - 1
        x = float(xl)
    else:
        xh = xl + 1
    t32 = lambda v: torch.tensor(v, dtype=torch.float32)  # noqa: E731
    ly, lx = t32(y) - yl, t32(x) - xl
    hy, hx = 1.0 - ly, 1.0 - lx
    return (hy * hx) * f[yl, xl] + (hy * lx) * f[yl, xh] + (ly * hx) * f[yh, xl] + (ly * lx) * f[yh, xh]


def mask_pool(fmap: Tensor, masks: Sequence[Tensor]) -> Tensor:
    """sum_p f_p m_p / (sum_p m_p + 1e-12) per mask, without materialising copies (E:645-653).

    fmap [B,h,w,C] (already unit-norm per token); masks: per-image [n,h,w]."""
    B, h, w, C = fmap.shape
    flat = fmap.reshape(B, h * w, C)
    out = []
    for b, m in enumerate(masks):
        m = m.float().flatten(-2, -1)                       # [n, hw]
        out.append((m @ flat[b]) / (m.sum(1, keepdim=True) + 1e-12))
    return torch.cat(out) if out else fmap.new_zeros(0, C)


def cosine_loss(student: Tensor, teacher: Tensor, weight: float = 1.0) -> Tensor:
    """(1 - mean_r <s/|s|, t/|t|>) * weight, eps 1e-12 (C:42-47)."""
    s = F.normalize(student, dim=-1)
    t = F.normalize(teacher, dim=-1)
    return (1.0 - (s * t).sum(-1).mean()) * weight


# --------------------------------------------------------------------------------------
# Whole step (C:7-49)
# --------------------------------------------------------------------------------------
def visual_sd(state_dict: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """Strip the 'visual.' prefix of a CustomCLIP state_dict."""
    return {k[len("visual."):]: v for k, v in state_dict.items() if k.startswith("visual.")}


def clipself_step(student_sd: Dict[str, Tensor], teacher_sd: Dict[str, Tensor],
                  images: Tensor, normed_boxes: Tensor, image_crops: Tensor, cfg: TowerCfg,
                  cosine_weight: float = 1.0) -> Dict[str, Tensor]:
    """One distillation forward (+ autograd graph if student_sd tensors require grad)."""
    rois, crop_idx = extract_rois(normed_boxes)
    crops = image_crops.flatten(0, 1)[crop_idx]
    with torch.no_grad():
        teacher = tower_forward_cls(teacher_sd, crops, cfg)
    dense = tower_encode_dense(student_sd, images, cfg)
    boxes = denormalize_boxes(rois, dense.shape[1], dense.shape[2])
    student = roi_align_1x1_nhwc(dense, boxes)
    loss = cosine_loss(student, teacher, cosine_weight)
    return dict(loss=loss, student_roi=student, teacher=teacher, dense=dense,
                crop_index=crop_idx, rois=torch.cat(boxes) if boxes else None)


# --------------------------------------------------------------------------------------
# Deterministic synthetic weights / batches shared by golden generation, tests and bench
# --------------------------------------------------------------------------------------
def tower_param_shapes(cfg: TowerCfg) -> List[Tuple[str, Tuple[int, ...]]]:
    """visual.* parameter names and shapes in state_dict order."""
    D, Hd, C, N = cfg.width, cfg.hidden, cfg.embed_dim, cfg.tokens
    out: List[Tuple[str, Tuple[int, ...]]] = [
        ("cls_token", (1, 1, D)), ("pos_embed", (1, N, D)),
        ("patch_embed.proj.weight", (D, 3, cfg.patch, cfg.patch)), ("patch_embed.proj.bias", (D,)),
    ]
    for i in range(cfg.layers):
        p = f"blocks.{i}."
        out += [(p + "norm1.weight", (D,)), (p + "norm1.bias", (D,)),
                (p + "attn.q_bias", (D,)), (p + "attn.v_bias", (D,)),
                (p + "attn.q_proj.weight", (D, D)), (p + "attn.k_proj.weight", (D, D)),
                (p + "attn.v_proj.weight", (D, D)),
                (p + "attn.inner_attn_ln.weight", (D,)), (p + "attn.inner_attn_ln.bias", (D,)),
                (p + "attn.proj.weight", (D, D)), (p + "attn.proj.bias", (D,)),
                (p + "norm2.weight", (D,)), (p + "norm2.bias", (D,)),
                (p + "mlp.w1.weight", (Hd, D)), (p + "mlp.w1.bias", (Hd,)),
                (p + "mlp.w2.weight", (Hd, D)), (p + "mlp.w2.bias", (Hd,)),
                (p + "mlp.ffn_ln.weight", (Hd,)), (p + "mlp.ffn_ln.bias", (Hd,)),
                (p + "mlp.w3.weight", (D, Hd)), (p + "mlp.w3.bias", (D,))]
    out += [("norm.weight", (D,)), ("norm.bias", (D,)),
            ("head.weight", (C, D)), ("head.bias", (C,))]
    return out


def synth_tower_weights(cfg: TowerCfg, seed: int) -> Dict[str, Tensor]:
    """Seeded synthetic tower weights (numpy PCG64 -> version independent).

    Not the reference initialiser on purpose: biases / LN affine / head are given
    non-trivial values so that every term of the path is exercised by parity tests
    (the reference init zeroes all biases and scales the head by 1e-3, E:455-467)."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: Dict[str, Tensor] = {}
    for name, shape in tower_param_shapes(cfg):
        if name.endswith("norm1.weight") or name.endswith("norm2.weight") or \
                name.endswith("ln.weight") or name == "norm.weight":
            a = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name.endswith(".bias") or name.endswith("_bias"):
            a = 0.05 * rng.standard_normal(shape)
        elif name in ("cls_token", "pos_embed"):
            a = 0.02 * rng.standard_normal(shape)
        elif name == "head.weight":
            a = 0.05 * rng.standard_normal(shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            a = rng.standard_normal(shape) * (0.7 / math.sqrt(fan_in))
        sd[name] = torch.from_numpy(a.astype(np.float32))
    return sd


def grid_box_templates(m: int, n: int) -> Tensor:
    """Row-major m x n grid boxes [x0,y0,x1,y1] from fp32 linspace (training/data.py:200-224)."""
    ys = torch.linspace(0, 1, m + 1)
    xs = torch.linspace(0, 1, n + 1)
    boxes = []
    for i in range(m):
        for j in range(n):
            boxes.append(torch.stack([xs[j], ys[i], xs[j + 1], ys[i + 1]]))
    return torch.stack(boxes)


def synth_batch(cfg: TowerCfg, batch: int, boxes_per_image: int, seed: int,
                kind: str = "grid", ragged: bool = False, crop_size: int | None = None,
                det_size: int | None = None) -> Tuple[Tensor, Tensor, Tensor]:
    """(images [B,3,S,S], normed_boxes [B,K,5], image_crops [B,K,3,s,s]) as the reference's
    datasets emit them (training/data.py:281), from a seeded numpy PCG64 stream.

    kind="grid": K boxes drawn without replacement from the smallest square grid template with
    at least K cells (training/data.py:200-232); kind="proposal": x0,y0~U(0,.6), w,h~U(.1,.4).
    ragged=True zeroes a random tail of each image's rows (valid flag 0, data.py:265,276-277)."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    S = det_size or cfg.image_size          # student (detector-resolution) image size, --det-image-size
    s = crop_size or cfg.image_size
    K = boxes_per_image
    images = torch.from_numpy(rng.standard_normal((batch, 3, S, S), dtype=np.float32))
    crops = torch.from_numpy(rng.standard_normal((batch, K, 3, s, s), dtype=np.float32))
    boxes = torch.zeros(batch, K, 5)
    if kind == "grid":
        side = 1
        while side * side < K:
            side += 1
        tmpl = grid_box_templates(side, side)
        for b in range(batch):
            perm = torch.from_numpy(rng.permutation(tmpl.shape[0])[:K].copy())
            boxes[b, :, :4] = tmpl[perm]
    elif kind == "proposal":
        xy = torch.from_numpy(rng.random((batch, K, 2), dtype=np.float32)) * 0.6
        wh = torch.from_numpy(rng.random((batch, K, 2), dtype=np.float32)) * 0.3 + 0.1
        boxes[..., 0:2] = xy
        boxes[..., 2:4] = xy + wh
    else:
        raise ValueError(kind)
    boxes[..., 4] = 1.0
    if ragged:
        keep = rng.integers(1, K + 1, size=batch)
        for b in range(batch):
            boxes[b, int(keep[b]):, :] = 0.0
    return images, boxes, crops
