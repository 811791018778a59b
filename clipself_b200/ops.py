"""Tensor-level wrappers over the C ABI (torch is only the allocator / stream provider here)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L
from ._lib import call

Tensor = torch.Tensor


# bench.py sets this to a list to time every GEMM launch with CUDA events on the launching stream
GEMM_PROFILE = None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _dt(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return L.CS_F32
    if t.dtype == torch.bfloat16:
        return L.CS_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _chk(t: Tensor, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise L.ClipselfB200Error(f"{name} must be a CUDA tensor (no CPU path exists)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


# ------------------------------------------------------------------ region path
def extract_rois(normed_boxes: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Device-side index extraction (clipself.py:29-36). Returns (rois[B*K,4], crop_index[B*K],
    roi_batch[B*K], img_offsets[B+1]); the first img_offsets[-1] rows are valid."""
    _chk(normed_boxes, torch.float32, "normed_boxes")
    B, K, five = normed_boxes.shape
    assert five == 5
    dev = normed_boxes.device
    rois = torch.empty(B * K, 4, device=dev, dtype=torch.float32)       # rows past the valid count are zeroed by the kernel
    crop_index = torch.empty(B * K, device=dev, dtype=torch.int32)
    roi_batch = torch.empty(B * K, device=dev, dtype=torch.int32)
    offsets = torch.empty(B + 1, device=dev, dtype=torch.int32)
    call("cs_extract_rois", _p(normed_boxes), B, K, _p(rois), _p(crop_index), _p(roi_batch), _p(offsets), _stream())
    return rois, crop_index, roi_batch, offsets


def gather_rows(src: Tensor, index: Tensor, R: int) -> Tensor:
    _chk(src, None, "src")
    _chk(index, torch.int32, "index")
    row_bytes = src[0].numel() * src.element_size()
    out = torch.empty((R,) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
    call("cs_gather_rows", _p(src), _p(index), R, row_bytes, _p(out), _stream())
    return out


def roi_align_fwd(fmap: Tensor, rois: Tensor, img_offsets: Tensor, R: int) -> Tuple[Tensor, Tensor, Tensor]:
    _chk(fmap, torch.float32, "fmap")
    B, H, W, Cc = fmap.shape
    out = torch.empty(R, Cc, device=fmap.device, dtype=torch.float32)
    wy = torch.empty(max(R, 1), H, device=fmap.device, dtype=torch.float32)
    wx = torch.empty(max(R, 1), W, device=fmap.device, dtype=torch.float32)
    call("cs_roi_align_fwd", _p(fmap), B, H, W, Cc, _p(rois), _p(img_offsets), R, _p(wy), _p(wx), _p(out), _stream())
    return out, wy, wx


def roi_align_bwd(d_out: Tensor, shape, img_offsets: Tensor, R: int, wy: Tensor, wx: Tensor) -> Tensor:
    _chk(d_out, torch.float32, "d_out")
    B, H, W, Cc = shape
    d_fmap = torch.empty(B, H, W, Cc, device=d_out.device, dtype=torch.float32)
    call("cs_roi_align_bwd", _p(d_out), B, H, W, Cc, _p(img_offsets), R, _p(wy), _p(wx), _p(d_fmap), _stream())
    return d_fmap


def mask_pool_fwd(fmap: Tensor, masks: Tensor, img_offsets: Tensor) -> Tensor:
    _chk(fmap, torch.float32, "fmap")
    _chk(masks, torch.float32, "masks")
    B, HW, Cc = fmap.shape
    R = masks.shape[0]
    out = torch.empty(R, Cc, device=fmap.device, dtype=torch.float32)
    call("cs_mask_pool_fwd", _p(fmap), B, HW, Cc, _p(masks), _p(img_offsets), R, _p(out), _stream())
    return out


def cosine_loss_fwd(s: Tensor, t: Tensor, weight: float) -> Tuple[Tensor, Tensor]:
    _chk(s, torch.float32, "s")
    _chk(t, torch.float32, "t")
    R, Cc = s.shape
    loss = torch.empty((), device=s.device, dtype=torch.float32)
    stats = torch.empty(R, 3, device=s.device, dtype=torch.float32)
    call("cs_cosine_loss_fwd", _p(s), _p(t), R, Cc, float(weight), _p(loss), _p(stats), _stream())
    return loss, stats


def cosine_loss_bwd(s: Tensor, t: Tensor, stats: Tensor, weight: float, d_loss: Tensor) -> Tensor:
    R, Cc = s.shape
    _chk(d_loss, torch.float32, "d_loss")
    d_s = torch.empty_like(s)
    call("cs_cosine_loss_bwd", _p(s), _p(t), _p(stats), R, Cc, float(weight), _p(d_loss), _p(d_s), _stream())
    return d_s


def l2norm_fwd(x: Tensor) -> Tuple[Tensor, Tensor]:
    _chk(x, torch.float32, "x")
    M, Cc = x.shape
    y = torch.empty_like(x)
    inv = torch.empty(M, device=x.device, dtype=torch.float32)
    call("cs_l2norm_fwd", _p(x), M, Cc, _p(y), _p(inv), _stream())
    return y, inv


def l2norm_bwd(y: Tensor, inv: Tensor, d_y: Tensor) -> Tensor:
    M, Cc = y.shape
    d_x = torch.empty_like(y)
    call("cs_l2norm_bwd", _p(y), _p(inv), _p(d_y), M, Cc, _p(d_x), _stream())
    return d_x


# ------------------------------------------------------------------ tower kernels
def im2col_patches(images: Tensor, patch: int, ldp: int) -> Tensor:
    _chk(images, None, "images")
    B, three, S, S2 = images.shape
    assert three == 3 and S == S2
    g = S // patch
    out = torch.empty(B * g * g, ldp, device=images.device, dtype=torch.bfloat16)
    call("cs_im2col_patches", _p(images), _dt(images), B, S, patch, _p(out), ldp, _stream())
    return out


def resize_bilinear(images: Tensor, size: int) -> Tensor:
    """F.interpolate(images, size=(size, size), mode='bilinear') for [B,C,H,W] f32 / bf16 (clipself.py:27)."""
    _chk(images, None, "images")
    B, Cc, H, W = images.shape
    out = torch.empty(B, Cc, size, size, device=images.device, dtype=images.dtype)
    call("cs_resize_bilinear", _p(images), _dt(images), B * Cc, H, W, size, size, _p(out), _stream())
    return out


def fill_cls_rows(cls_token: Tensor, pos_embed: Tensor, x: Tensor) -> None:
    B, N, D = x.shape
    call("cs_fill_cls_rows", _p(cls_token), _p(pos_embed), B, N, D, _p(x), _stream())


def layernorm_fwd(x: Tensor, M: int, D: int, gamma: Tensor, beta: Tensor, eps: float, out: Tensor,
                  ldx: Optional[int] = None, row_div: int = 0, row_mul: int = 1, row_off: int = 0,
                  mean: Optional[Tensor] = None, rstd: Optional[Tensor] = None) -> Tensor:
    ldx = ldx if ldx is not None else x.shape[-1]
    call("cs_layernorm_fwd", _p(x), _dt(x), ldx, M, D, row_div, row_mul, row_off, _p(gamma), _p(beta),
         float(eps), _p(out), out.shape[-1], _p(mean), _p(rstd), _stream())
    return out


def row_stats_cast(x: Tensor, M: int, D: int, xb: Tensor, stats: Tensor) -> None:
    """f32 rows -> bf16 copy + LayerNorm partial statistics [M, parts, 2] (whole row in part 0)."""
    _chk(x, torch.float32, "x")
    _chk(xb, torch.bfloat16, "xb")
    call("cs_row_stats_cast", _p(x), x.shape[-1], M, D, _p(xb), xb.shape[-1], _p(stats), stats.shape[-2], _stream())


def gemm(a: Tensor, w: Tensor, out: Tensor, *, M: Optional[int] = None, N: Optional[int] = None,
         K: Optional[int] = None, mode: int = L.EPI_STORE, bias: Optional[Tensor] = None,
         residual: Optional[Tensor] = None, rope: Optional[Tuple[Tensor, Tensor]] = None, tokens: int = 0,
         rope_cols: int = 0, pos_embed: Optional[Tensor] = None, alpha: float = 1.0,
         ldo: Optional[int] = None, dbg: int = 0, ln_fold: Optional[Tuple[Tensor, Tensor, int, int, float]] = None,
         stats_out: Optional[Tensor] = None, k_splits: int = 0, out2: Optional[Tensor] = None) -> Tensor:
    """out = epilogue(a[M,K] @ w[N,K]^T); a, w bf16 row-major (lda/ldw = last dim)."""
    _chk(a, torch.bfloat16, "a")
    _chk(w, torch.bfloat16, "w")
    M = M if M is not None else a.shape[0]
    N = N if N is not None else w.shape[0]
    K = K if K is not None else a.shape[1]
    e = L.GemmEpilogue()
    e.mode = mode
    e.out_dtype = _dt(out)
    e.out = _p(out)
    e.ldo = ldo if ldo is not None else out.shape[-1]
    e.bias = _p(bias)
    e.residual = _p(residual)
    e.ldr = residual.shape[-1] if residual is not None else 0
    e.rope_pos = _p(rope[0]) if rope is not None else None      # rope = (pos [grid], freq [16])
    e.rope_freq = _p(rope[1]) if rope is not None else None
    e.rope_grid = rope[0].numel() if rope is not None else 0
    e.tokens = tokens
    e.rope_cols = rope_cols
    e.pos_embed = _p(pos_embed)
    e.alpha = alpha
    e.reserved = dbg
    if ln_fold is not None:          # (stats [M,parts,2], c1 [N], parts, dim, eps)
        e.ln_stats, e.ln_c1 = _p(ln_fold[0]), _p(ln_fold[1])
        e.ln_parts, e.ln_dim, e.ln_eps = int(ln_fold[2]), int(ln_fold[3]), float(ln_fold[4])
    e.stats_out = _p(stats_out)
    if out2 is not None:             # bf16 copy of the f32 output (+ stats_out: its row statistics)
        e.out2_bf16, e.ldo2 = _p(out2), out2.shape[-1]
    e.reserved2 = k_splits           # split-K (out must be pre-zeroed): -1 = let the library choose
    if GEMM_PROFILE is None:
        call("cs_gemm_bf16", _p(a), a.shape[-1], _p(w), w.shape[-1], M, N, K, C.byref(e), _stream())
        return out
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call("cs_gemm_bf16", _p(a), a.shape[-1], _p(w), w.shape[-1], M, N, K, C.byref(e), _stream())
    e1.record()
    GEMM_PROFILE.append((2.0 * M * N * K, e0, e1))
    return out


def gemm_tn(at: Tensor, bt: Tensor, out: Tensor, *, M: int, N: int, K: int, lda: Optional[int] = None,
            ldb: Optional[int] = None, ldo: Optional[int] = None, k_splits: int = 0) -> Tensor:
    """out[M,N] (f32) = at[K,M]^T @ bt[K,N]: the weight-gradient GEMM on operands stored tokens-major (no transposes).
    k_splits as in gemm() (out must be pre-zeroed when != 0)."""
    for name, x in (("at", at), ("bt", bt)):        # column-sliced views are fine: the leading dimension is explicit
        if not x.is_cuda:
            raise L.ClipselfB200Error(f"{name} must be a CUDA tensor (no CPU path exists)")
        if x.dtype != torch.bfloat16:
            raise TypeError(f"{name}: expected bf16, got {x.dtype}")
    e = L.GemmEpilogue()
    e.mode = L.EPI_STORE
    e.out_dtype = _dt(out)
    e.out = _p(out)
    e.ldo = ldo if ldo is not None else out.shape[-1]
    e.alpha = 1.0
    e.reserved2 = k_splits
    args = (_p(at), lda if lda is not None else at.stride(0), _p(bt), ldb if ldb is not None else bt.stride(0), M, N, K,
            C.byref(e), _stream())
    if GEMM_PROFILE is None:
        call("cs_gemm_bf16_tn", *args)
        return out
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call("cs_gemm_bf16_tn", *args)
    e1.record()
    GEMM_PROFILE.append((2.0 * M * N * K, e0, e1))
    return out


def pack_swiglu_weights(w1: Tensor, w2: Tensor, b1: Tensor, b2: Tensor, ldk: int) -> Tuple[Tensor, Tensor]:
    Hd, K = w1.shape
    rows = (Hd + 127) // 128 * 256
    packed = torch.empty(rows, ldk, device=w1.device, dtype=torch.bfloat16)
    bias = torch.empty(rows, device=w1.device, dtype=torch.float32)
    call("cs_pack_swiglu_weights", _p(w1), _p(w2), _dt(w1), Hd, K, _p(b1), _p(b2), _p(packed), ldk, _p(bias), _stream())
    return packed, bias


def attention_fwd(qkv: Tensor, B: int, N: int, H: int, scale: float, out: Tensor,
                  lse: Optional[Tensor] = None, row_stats: Optional[Tensor] = None) -> Tensor:
    _chk(qkv, torch.bfloat16, "qkv")
    call("cs_attention_fwd", _p(qkv), B, N, H, float(scale), _p(out), _p(lse), _p(row_stats), _stream())
    return out


def attention_cls_fwd(qkv: Tensor, B: int, N: int, H: int, scale: float, out_cls: Tensor,
                      row_stats_cls: Optional[Tensor] = None) -> Tensor:
    """softmax(q_cls k^T * scale) v for the CLS query of every (image, head): qkv [B*N, 3*H*64] -> out_cls [B, H*64]."""
    _chk(qkv, torch.bfloat16, "qkv")
    call("cs_attention_cls_fwd", _p(qkv), B, N, H, float(scale), _p(out_cls), _p(row_stats_cls), _stream())
    return out_cls


def cast_pad_bf16(src: Tensor, ldd: Optional[int] = None) -> Tensor:
    _chk(src, torch.float32, "src")
    src2 = src.reshape(src.shape[0], -1)
    rows, cols = src2.shape
    ldd = ldd if ldd is not None else (cols + 7) // 8 * 8
    out = torch.empty(rows, ldd, device=src.device, dtype=torch.bfloat16)
    call("cs_cast_pad_bf16", _p(src2), rows, cols, cols, _p(out), ldd, _stream())
    return out


# ------------------------------------------------------------------ backward kernels
def attention_bwd(qkv: Tensor, out: Tensor, d_out: Tensor, lse: Tensor, B: int, N: int, H: int, scale: float,
                  rope: Optional[Tuple[Tensor, Tensor]], delta_ws: Tensor, dqkv: Tensor) -> Tensor:
    call("cs_attention_bwd", _p(qkv), _p(out), _p(d_out), _p(lse), B, N, H, float(scale),
         _p(rope[0]) if rope else None, _p(rope[1]) if rope else None, _p(delta_ws), _p(dqkv), _stream())
    return dqkv


def cast_transpose(src: Tensor, M: int, N: int, dst: Optional[Tensor] = None, dst_t: Optional[Tensor] = None,
                   lds: Optional[int] = None, ldd: Optional[int] = None, ldt: Optional[int] = None) -> None:
    """src [M,N] (f32|bf16) -> dst [M,*] bf16 and/or dst_t [N,*] bf16 (ldd / ldt override the leading
    dimensions when dst / dst_t are column-sliced views of wider matrices)."""
    call("cs_cast_transpose_bf16", _p(src), _dt(src), M, N, lds if lds is not None else src.shape[-1],
         _p(dst), (ldd if ldd is not None else dst.shape[-1]) if dst is not None else 0, _p(dst_t),
         (ldt if ldt is not None else dst_t.shape[-1]) if dst_t is not None else 0, _stream())


def layernorm_bwd_dx(dy: Tensor, x: Tensor, M: int, D: int, mean: Tensor, rstd: Tensor, gamma: Tensor, dx: Tensor,
                     add: Optional[Tensor] = None, row_div: int = 0, row_mul: int = 1, row_off: int = 0,
                     row_mapped: bool = False) -> Tensor:
    call("cs_layernorm_bwd_dx", _p(dy), _dt(dy), dy.shape[-1], _p(x), _dt(x), x.shape[-1], M, D, row_div, row_mul,
         row_off, _p(mean), _p(rstd), _p(gamma), _p(add), add.shape[-1] if add is not None else 0, _p(dx), _dt(dx),
         dx.shape[-1], int(row_mapped), _stream())
    return dx


def col_reduce(dy: Tensor, M: int, D: int, dbeta: Tensor, workspace: Tensor, x: Optional[Tensor] = None,
               mean: Optional[Tensor] = None, rstd: Optional[Tensor] = None, dgamma: Optional[Tensor] = None,
               row_div: int = 0, row_mul: int = 1, row_off: int = 0, lddy: Optional[int] = None) -> None:
    call("cs_col_reduce", _p(dy), _dt(dy), lddy if lddy is not None else dy.shape[-1], _p(x),
         _dt(x) if x is not None else 0, x.shape[-1] if x is not None else 0, M, D, row_div, row_mul, row_off,
         _p(mean), _p(rstd), _p(dgamma), _p(dbeta), _p(workspace), workspace.numel(), _stream())


def swiglu_fwd(x12: Tensor, M: int, Hd: int, h: Tensor, split=True) -> Tensor:
    """split: False = packed128 layout, True = up columns at Hd, int >= 2 = up columns at that offset."""
    call("cs_swiglu_fwd", _p(x12), M, Hd, x12.shape[-1], _p(h), h.shape[-1], int(split), _stream())
    return h


def swiglu_bwd(x12: Tensor, dh: Tensor, M: int, Hd: int, dx12: Tensor, split=True) -> Tensor:
    call("cs_swiglu_bwd", _p(x12), _p(dh), M, Hd, x12.shape[-1], dh.shape[-1], _p(dx12), int(split), _stream())
    return dx12


def adamw_step(param: Tensor, grad: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, lr: float, beta1: float,
               beta2: float, eps: float, weight_decay: float, step: int, grad_scale: float = 1.0) -> None:
    n = param.numel()
    assert grad.numel() == n and exp_avg.numel() == n and exp_avg_sq.numel() == n
    call("cs_adamw_step", _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), n, float(lr), float(beta1), float(beta2),
         float(eps), float(weight_decay), int(step), float(grad_scale), _stream())
