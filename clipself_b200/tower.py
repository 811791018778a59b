"""EVA02 vision tower engine: packed weights + kernel sequencing for the CLIPSelf hot path.

Mirrors (reference file:line, wusize/CLIPSelf @ 1c7fe9c):
  teacher  : EVAVisionTransformer.forward / forward_features   eva_vit_model.py:533-586
  student  : EVAVisionTransformer.encode_dense                  eva_vit_model.py:588-623
  blocks   : Block.forward / forward_without_attn               eva_vit_model.py:300-324
  attention: Attention.forward / proj_without_attn              eva_vit_model.py:174-256
  mlp      : SwiGLU.forward                                     eva_vit_model.py:98-105
  rope     : VisionRotaryEmbeddingFast                          rope.py:96-164

Precision contract (the reference's runnable bf16 mode, `--precision amp_bf16`):
fp32 master weights and residual stream, bf16 tensor-core operands, fp32 accumulation,
fp32 LayerNorm / softmax / normalise / RoIAlign / loss.

All arithmetic runs in the CUDA library (clipself_b200/csrc); torch only owns the memory.
"""
from __future__ import annotations

import copy
import dataclasses
import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import ops

Tensor = torch.Tensor


@dataclass(frozen=True)
class TowerCfg:
    image_size: int = 224
    patch: int = 16
    width: int = 768
    heads: int = 12
    layers: int = 12
    hidden: int = 2048
    embed_dim: int = 512
    pt_seq_len: int = 16
    ln_eps: float = 1e-6

    @property
    def grid(self) -> int:
        return self.image_size // self.patch

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + 1

    @property
    def head_dim(self) -> int:
        return self.width // self.heads

    @property
    def hidden_pad(self) -> int:
        """SwiGLU hidden width padded to the 128-column packing unit (ViT-L: 2730 -> 2816); padded
        columns carry exact zeros (zero weights / biases) and LayerNorm statistics use `hidden`."""
        return (self.hidden + 127) // 128 * 128


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _attention_emits_stats(tokens: int) -> bool:
    """The LayerNorm row statistics come from the tcgen05 attention epilogues (both the N <= 224 kernel and the
    streaming one); only the mma.sync A/B kernel (CS_ATTN_LEGACY=1) has none."""
    return os.environ.get("CS_ATTN_LEGACY", "0") in ("", "0")


def stat_parts(n: int) -> int:
    """Partial (sum, sumsq) pairs per row that a GEMM epilogue writes for an n-column output: one per
    (n-tile, column half), tiles of 256 columns when n % 256 == 0, else 128 (cs_gemm_epilogue_t.stats_out)."""
    t = 256 if n % 256 == 0 else 128
    return 2 * ((n + t - 1) // t)


def chunk_schedule(rows: int, step: int) -> List[tuple]:
    """(start, count) pieces of a teacher pass over `rows` crops: a short first and second piece (step/4,
    step/2) so the pass can start as soon as the first few crops have crossed PCIe, then full `step`s."""
    out, s = [], 0
    for n in (max(step // 4, 1), max(step // 2, 1)):
        if rows - s > step:                  # only worth splitting when more than one full piece remains
            out.append((s, n))
            s += n
    while s < rows:
        n = min(step, rows - s)
        out.append((s, n))
        s += n
    return out


def input_grid(images: Tensor, cfg: "TowerCfg") -> int:
    """Token grid of an input batch. Like the reference's towers (eva_vit_model.py:533-549) any square
    resolution that is a multiple of the patch size is accepted; the grid is capped by the RoPE table the
    QKV epilogue keeps in shared memory (64 x 64 = 1024 px at /16, 896 px at /14: the published recipes)."""
    S = images.shape[-1]
    if images.shape[-2] != S or S % cfg.patch != 0 or not 0 < S // cfg.patch <= 64:
        raise ValueError(f"input must be square, a multiple of the patch size {cfg.patch} and at most "
                         f"{64 * cfg.patch} px, got {tuple(images.shape[-2:])}")
    return S // cfg.patch


def rescale_pos_embed(pos_embed: Tensor, grid: int) -> Tensor:
    """rescale_positional_embedding (eva_vit_model.py:631-643): the CLS row is kept, the [g0,g0] grid of
    patch rows is resampled bicubically (align_corners=False) to [grid,grid]. -> [1 + grid^2, D] f32.
    Parameter preprocessing, done once per resolution (pos_embed is frozen on this path)."""
    pos = pos_embed.detach().reshape(-1, pos_embed.shape[-1]).float()
    g0 = int(round(math.sqrt(pos.shape[0] - 1)))
    if g0 == grid:
        return pos.contiguous()
    pe = pos[1:].T.contiguous().view(1, -1, g0, g0)
    pe = torch.nn.functional.interpolate(pe, (grid, grid), mode="bicubic", align_corners=False).view(-1, grid * grid)
    return torch.cat([pos[:1], pe.T], dim=0).contiguous()


def rope_tables(grid: int, head_dim: int, pt_seq_len: int, theta: float = 10000.0):
    """cos/sin [grid*grid, head_dim] f32, built with the same torch ops / order as rope.py:118-142
    so the tables are bit-identical to the reference's registered buffers."""
    dim = head_dim // 2
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    t = torch.arange(grid) / grid * pt_seq_len
    ang = torch.einsum("..., f -> ... f", t, freqs)
    ang = ang.repeat_interleave(2, dim=-1)
    full = torch.cat([ang[:, None, :].expand(grid, grid, dim), ang[None, :, :].expand(grid, grid, dim)], dim=-1)
    return full.cos().reshape(-1, 2 * dim).contiguous(), full.sin().reshape(-1, 2 * dim).contiguous()


def rope_vectors(grid: int, head_dim: int, pt_seq_len: int, theta: float = 10000.0):
    """The two 1-D factors of the tables above: pos [grid] (rope.py:127) and freq [head_dim/4]
    (rope.py:118), same torch ops -> angle(token, d) = pos[row or col] * freq[(d % 32) // 2]."""
    dim = head_dim // 2
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    t = torch.arange(grid) / grid * pt_seq_len
    return t.float().contiguous(), freqs.contiguous()


class PackedBlock:
    __slots__ = ("wqkv", "bqkv", "wv", "bv", "wproj", "bproj", "w12", "b12", "w3", "b3",
                 "g1", "b1", "gi", "bi", "g2", "b2", "gf", "bf",
                 # LayerNorm-folded operands (frozen weights): W*diag(gamma), rowsum(W'), W*beta + b
                 "wproj_f", "c1_proj", "c2_proj", "w3_f", "c1_w3", "c2_w3",
                 "wqkv_f", "c1_qkv", "c2_qkv", "w12_f", "c1_w12", "c2_w12")


class PackedTower:
    """GEMM-ready bf16 copies of one tower's weights (rebuilt after every optimizer step for
    the student, once for the frozen teacher)."""

    def __init__(self, cfg: TowerCfg, sd: Dict[str, Tensor], device: torch.device):
        assert cfg.head_dim == 64, "kernels are specialised for head_dim 64 (all EVA02-CLIP configs)"
        assert cfg.width % 64 == 0 and cfg.embed_dim % 32 == 0
        self.cfg = cfg
        self.device = device
        cos, sin = rope_tables(cfg.grid, cfg.head_dim, cfg.pt_seq_len)
        self.rope_cos = cos.to(device)
        self.rope_sin = sin.to(device)
        pos, freq = rope_vectors(cfg.grid, cfg.head_dim, cfg.pt_seq_len)
        self.rope_pos, self.rope_freq = pos.to(device), freq.to(device)
        self.k_pe = 3 * cfg.patch * cfg.patch
        self.k_pe_pad = _round_up(self.k_pe, 8)
        self.blocks: List[PackedBlock] = [PackedBlock() for _ in range(cfg.layers)]
        self.repack(sd)

    @staticmethod
    def _f32(t: Tensor, device) -> Tensor:
        return t.detach().to(device=device, dtype=torch.float32).contiguous()

    def repack(self, sd: Dict[str, Tensor]) -> None:
        cfg, dev = self.cfg, self.device
        f = lambda k: self._f32(sd[k], dev)  # noqa: E731
        D = cfg.width
        self.pe_w = ops.cast_pad_bf16(f("patch_embed.proj.weight").reshape(D, -1), self.k_pe_pad)
        self.pe_b = f("patch_embed.proj.bias")
        self.cls = f("cls_token").reshape(-1)
        self.pos_src = f("pos_embed")
        self.pos = rescale_pos_embed(self.pos_src, cfg.grid)
        self.norm_g, self.norm_b = f("norm.weight"), f("norm.bias")
        self.head_w = ops.cast_pad_bf16(f("head.weight"))
        self.head_b = f("head.bias")
        for i, pb in enumerate(self.blocks):
            p = f"blocks.{i}."
            wq, wk, wv = f(p + "attn.q_proj.weight"), f(p + "attn.k_proj.weight"), f(p + "attn.v_proj.weight")
            pb.wqkv = ops.cast_pad_bf16(torch.cat([wq, wk, wv], dim=0))
            pb.bqkv = torch.cat([f(p + "attn.q_bias"), torch.zeros(D, device=dev), f(p + "attn.v_bias")])
            pb.wv = pb.wqkv[2 * D:]
            pb.bv = pb.bqkv[2 * D:]
            pb.wproj = ops.cast_pad_bf16(f(p + "attn.proj.weight"))
            pb.bproj = f(p + "attn.proj.bias")
            pb.w12, pb.b12 = ops.pack_swiglu_weights(f(p + "mlp.w1.weight"), f(p + "mlp.w2.weight"),
                                                     f(p + "mlp.w1.bias"), f(p + "mlp.w2.bias"), D)
            pb.w3 = ops.cast_pad_bf16(f(p + "mlp.w3.weight"), cfg.hidden_pad)
            pb.b3 = f(p + "mlp.w3.bias")
            pb.g1, pb.b1 = f(p + "norm1.weight"), f(p + "norm1.bias")
            pb.gi, pb.bi = f(p + "attn.inner_attn_ln.weight"), f(p + "attn.inner_attn_ln.bias")
            pb.g2, pb.b2 = f(p + "norm2.weight"), f(p + "norm2.bias")
            pb.gf, pb.bf = f(p + "mlp.ffn_ln.weight"), f(p + "mlp.ffn_ln.bias")
            # One-time weight preprocessing for the folded inner_attn_ln / ffn_ln (see cs_gemm_epilogue_t):
            #   proj(LN(a)) = rstd * (a @ W'^T - mean * c1) + c2,  W' = W diag(gamma), c1 = rowsum(W'), c2 = W beta + b.
            # c1 is summed from the bf16-rounded W' so it matches what the tensor cores multiply.
            wp, w3 = f(p + "attn.proj.weight"), f(p + "mlp.w3.weight")
            pb.wproj_f = ops.cast_pad_bf16(wp * pb.gi[None, :])
            pb.c1_proj = pb.wproj_f.float().sum(1).contiguous()
            pb.c2_proj = (wp @ pb.bi + pb.bproj).contiguous()
            pb.w3_f = ops.cast_pad_bf16(w3 * pb.gf[None, :], cfg.hidden_pad)
            pb.c1_w3 = pb.w3_f.float().sum(1).contiguous()
            pb.c2_w3 = (w3 @ pb.bf + pb.b3).contiguous()
            # norm1 folded into q|k|v, norm2 into w1|w2 (same algebra; the GEMMs then read the bf16 copy of the
            # residual stream that the previous block's epilogue wrote, and no LayerNorm pass exists at all)
            wqkv = torch.cat([wq, wk, wv], dim=0)
            pb.wqkv_f = ops.cast_pad_bf16(wqkv * pb.g1[None, :])
            pb.c1_qkv = pb.wqkv_f.float().sum(1).contiguous()
            pb.c2_qkv = (wqkv @ pb.b1 + pb.bqkv).contiguous()
            w1, w2 = f(p + "mlp.w1.weight"), f(p + "mlp.w2.weight")
            pb.w12_f, pb.c2_w12 = ops.pack_swiglu_weights(w1 * pb.g2[None, :], w2 * pb.g2[None, :],
                                                          w1 @ pb.b2 + f(p + "mlp.w1.bias"), w2 @ pb.b2 + f(p + "mlp.w2.bias"), D)
            pb.c1_w12 = pb.w12_f.float().sum(1).contiguous()


class NativeTower:
    """Owner of a cs_tower_t handle (include/clipself_b200.h, tower level) and of the device memory it points into."""

    def __init__(self, cfg: TowerCfg, sd: Dict[str, Tensor], device: torch.device):
        import ctypes as C
        self.cfg, self.device = cfg, device
        self.ccfg = L.TowerCfgC(cfg.image_size, cfg.patch, cfg.width, cfg.heads, cfg.layers, cfg.hidden, cfg.embed_dim,
                                cfg.pt_seq_len, cfg.ln_eps)
        need = C.c_int64(0)
        L.call("cs_pack_weights_bytes", C.byref(self.ccfg), C.byref(need))
        self.pack = torch.empty(int(need.value), device=device, dtype=torch.uint8)
        self.handle = C.c_void_p()
        names, ptrs, self._keep = self._tensor_table(sd)
        L.call("cs_pack_weights_create", C.byref(self.ccfg), names, ptrs, len(self._keep), self.pack.data_ptr(), int(need.value),
               torch.cuda.current_stream().cuda_stream, C.byref(self.handle))
        self._keep = None
        self._ws: Optional[Tensor] = None
        self._ws_key = None
        self.launches_per_chunk = 5 * cfg.layers + 6

    def _tensor_table(self, sd: Dict[str, Tensor]):
        import ctypes as C
        keep = [(k, v.detach().to(device=self.device, dtype=torch.float32).contiguous()) for k, v in sd.items() if "rope" not in k]
        names = (C.c_char_p * len(keep))(*[k.encode() for k, _ in keep])
        ptrs = (C.c_void_p * len(keep))(*[t.data_ptr() for _, t in keep])
        return names, ptrs, keep

    def update(self, sd: Dict[str, Tensor]) -> None:
        names, ptrs, keep = self._tensor_table(sd)
        L.call("cs_pack_weights_update", self.handle, names, ptrs, len(keep), torch.cuda.current_stream().cuda_stream)
        torch.cuda.current_stream().synchronize()       # `keep` (temporary f32 copies) may be freed after this

    def workspace(self, chunk: int, image_size: int) -> Tensor:
        import ctypes as C
        key = (chunk, image_size)
        if self._ws_key is None or self._ws_key != key:
            need = C.c_int64(0)
            L.call("cs_query_workspace", C.byref(self.ccfg), chunk, image_size, C.byref(need))
            if self._ws is None or self._ws.numel() < need.value:
                self._ws = None
                self._ws = torch.empty(int(need.value), device=self.device, dtype=torch.uint8)
            self._ws_key = key
        return self._ws

    def forward_cls(self, images: Tensor, out: Tensor, chunk: int) -> None:
        assert images.is_contiguous() and out.is_contiguous() and out.dtype == torch.float32
        n = images.shape[0]
        ws = self.workspace(chunk, self.cfg.image_size)
        L.call("cs_vit_forward_cls", self.handle, images.data_ptr(), ops._dt(images), n, ws.data_ptr(), ws.numel(), chunk,
               out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        L.launch_count += self.launches_per_chunk * ((n + chunk - 1) // chunk)

    def forward_dense(self, images: Tensor, out: Tensor, chunk: int, image_size: int, pos: Optional[Tensor]) -> None:
        assert images.is_contiguous() and out.is_contiguous() and out.dtype == torch.float32
        n = images.shape[0]
        ws = self.workspace(chunk, image_size)
        L.call("cs_vit_forward_dense", self.handle, images.data_ptr(), ops._dt(images), n, image_size,
               pos.data_ptr() if pos is not None else None, ws.data_ptr(), ws.numel(), chunk, out.data_ptr(),
               torch.cuda.current_stream().cuda_stream)
        L.launch_count += (self.launches_per_chunk + 3) * ((n + chunk - 1) // chunk)

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                L.lib().cs_pack_weights_destroy(self.handle)
        except Exception:           # noqa: BLE001  (interpreter shutdown)
            pass


class Workspace:
    """Activation scratch for one forward-only chunk of `rows` token rows."""

    def __init__(self, cfg: TowerCfg, rows: int, device):
        D, Hd = cfg.width, cfg.hidden_pad
        bf = dict(device=device, dtype=torch.bfloat16)
        self.rows = rows
        self.x = torch.empty(rows, D, device=device, dtype=torch.float32)
        self.u = torch.empty(rows, D, **bf)
        self.qkv = torch.empty(rows, 3 * D, **bf)
        self.att = torch.empty(rows, D, **bf)
        self.h = torch.empty(rows, Hd, **bf)
        self.h2 = torch.zeros(rows, Hd, **bf)           # padded columns must stay finite (they meet zero weights)
        self.stats_att = torch.empty(rows, 4 * cfg.heads, 2, device=device, dtype=torch.float32)
        self.stats_h = torch.empty(rows, Hd // 64, 2, device=device, dtype=torch.float32)
        # bf16 copy of the residual stream + its row statistics (written by the proj / w3 / embed epilogues)
        self.xb = torch.empty(rows, D, **bf)
        self.stats_x = torch.empty(rows, stat_parts(D), 2, device=device, dtype=torch.float32)


class TowerEngine:
    """Runs the kernel sequence of one tower."""

    def __init__(self, cfg: TowerCfg, sd: Dict[str, Tensor], device: torch.device, chunk_images: Optional[int] = None):
        L.require_device()
        if chunk_images is None:
            # ~100k token rows per pass for the 197-token towers (measured: 512 crops 538 img/s, 384: 534, 256: 529,
            # 128: 455 on cfg2); the 577-token ViT-L keeps 256 crops (148k rows)
            chunk_images = int(os.environ.get("CLIPSELF_TEACHER_CHUNK", "512" if cfg.tokens <= 224 else "256"))
        self.cfg = cfg
        self.device = device
        self.w = PackedTower(cfg, sd, device)
        self.chunk_images = chunk_images
        self._ws: Optional[Workspace] = None
        self.scale = cfg.head_dim ** -0.5
        # LayerNorm folding needs the producers' row statistics: the tcgen05 attention kernel (N <= 224)
        # and an even number of SwiGLU tiles per row
        self.fold_proj = _attention_emits_stats(cfg.tokens) and os.environ.get("CLIPSELF_NO_LN_FOLD") is None
        self.fold_w3 = os.environ.get("CLIPSELF_NO_LN_FOLD") is None
        # norm1 / norm2 folded as well (CLIPSELF_NO_NORM_FOLD=1 keeps them as explicit LayerNorm kernels for A/B)
        self.fold_norm = self.fold_proj and self.fold_w3 and os.environ.get("CLIPSELF_NO_NORM_FOLD") is None

        self._views: Dict[int, "TowerEngine"] = {}
        self._graphs: Dict[tuple, object] = {}
        self._cls_ln: Optional[Tensor] = None
        self.use_graphs = os.environ.get("CLIPSELF_NO_GRAPH") is None
        # The product path: the tower-level C ABI (csrc/tower.cu: cs_pack_weights_* / cs_vit_forward_*) packs, sequences and
        # graph-replays the folded pipeline natively.  The Python sequencing above/below (block_inplace) is the same kernel
        # sequence kept for the stage-wise parity tests, per-GEMM event profiling and the un-folded A/B switches.
        self.native: Optional[NativeTower] = None
        if self.fold_norm and os.environ.get("CLIPSELF_PY_TOWER") is None:
            self.native = NativeTower(cfg, sd, device)

    def repack(self, sd: Dict[str, Tensor]) -> None:
        self.w.repack(sd)            # new packed tensors: captured graphs point at the old ones
        self._views.clear()
        self._graphs.clear()
        if self.native is not None:
            self.native.update(sd)

    def at_grid(self, grid: int) -> "TowerEngine":
        """The same tower at another input resolution (token grid): shares the packed block weights,
        owns its RoPE vectors, rescaled pos_embed and workspace (rope.py:179-214, eva_vit_model.py:631-643)."""
        if grid == self.cfg.grid:
            return self
        v = self._views.get(grid)
        if v is None:
            cfg = dataclasses.replace(self.cfg, image_size=grid * self.cfg.patch)
            v = TowerEngine.__new__(TowerEngine)
            v.cfg, v.device, v.scale = cfg, self.device, self.scale
            v.w = copy.copy(self.w)
            v.w.cfg = cfg
            cos, sin = rope_tables(grid, cfg.head_dim, cfg.pt_seq_len)
            v.w.rope_cos, v.w.rope_sin = cos.to(self.device), sin.to(self.device)
            pos, freq = rope_vectors(grid, cfg.head_dim, cfg.pt_seq_len)
            v.w.rope_pos, v.w.rope_freq = pos.to(self.device), freq.to(self.device)
            v.w.pos = rescale_pos_embed(self.w.pos_src, grid)
            v.chunk_images = max(1, self.chunk_images * self.cfg.tokens // cfg.tokens)
            v._ws = None
            v._graphs, v._cls_ln, v.use_graphs = {}, None, self.use_graphs
            v.native = self.native          # one handle serves every resolution (per-grid RoPE vectors are cached inside)
            v.fold_proj, v.fold_w3, v.fold_norm = self.fold_proj, self.fold_w3, self.fold_norm
            v._views = {}
            self._views[grid] = v
        return v

    # ------------------------------------------------------------------ helpers
    def workspace(self, images: int) -> Workspace:
        rows = images * self.cfg.tokens
        if self._ws is None or self._ws.rows < rows:
            self._graphs.clear()
            self._ws = Workspace(self.cfg, rows, self.device)
        return self._ws

    def embed(self, images: Tensor, x: Tensor, ws: Optional[Workspace] = None) -> None:
        """patch conv as a GEMM + bias + pos_embed, CLS rows (eva_vit_model.py:350-356, 540-544); with the fully folded
        pipeline also the bf16 copy + row statistics of x that the first block's QKV GEMM consumes."""
        cfg, w = self.cfg, self.w
        B = images.shape[0]
        patches = ops.im2col_patches(images, cfg.patch, w.k_pe_pad)
        ops.gemm(patches, w.pe_w, x, M=B * (cfg.tokens - 1), N=cfg.width, K=w.k_pe_pad, mode=L.EPI_TOKENS,
                 bias=w.pe_b, pos_embed=w.pos, tokens=cfg.tokens)
        ops.fill_cls_rows(w.cls, w.pos, x[:B * cfg.tokens].view(B, cfg.tokens, cfg.width))
        if ws is not None and self.fold_norm:
            ops.row_stats_cast(x, B * cfg.tokens, cfg.width, ws.xb, ws.stats_x)

    def block_inplace(self, i: int, ws: Workspace, B: int, with_attention: bool = True) -> None:
        """One residual block on ws.x in place (inference; nothing saved).

        Fully folded form (default), 5 launches and no LayerNorm pass:
            qkv  = rope(LN1-fold(xb Wqkv'^T))                       xb = bf16(x), statistics of x from the producer
            att  = softmax(q k^T / 8) v                             (+ row statistics of att)
            x   += LNi-fold(att Wproj'^T)        -> x, xb, stats_x  (one epilogue)
            h    = silu(.)*(.) of LN2-fold(xb W12'^T)               (+ row statistics of h)
            x   += LNf-fold(h W3'^T)             -> x, xb, stats_x
        """
        cfg, pb = self.cfg, self.w.blocks[i]
        D, N = cfg.width, cfg.tokens
        M = B * N
        x, u = ws.x, ws.u
        eps = cfg.ln_eps
        fold_proj = with_attention and self.fold_proj
        if self.fold_norm:
            sx = (ws.stats_x, stat_parts(D), D, eps)
            if with_attention:
                ops.gemm(ws.xb, pb.wqkv_f, ws.qkv, M=M, mode=L.EPI_QKV_ROPE, bias=pb.c2_qkv, rope=(self.w.rope_pos, self.w.rope_freq),
                         tokens=N, rope_cols=2 * D, ln_fold=(sx[0], pb.c1_qkv, *sx[1:]))
                ops.attention_fwd(ws.qkv, B, N, cfg.heads, self.scale, ws.att, row_stats=ws.stats_att)
                ops.gemm(ws.att, pb.wproj_f, x, M=M, bias=pb.c2_proj, residual=x, out2=ws.xb, stats_out=ws.stats_x,
                         ln_fold=(ws.stats_att, pb.c1_proj, 4 * cfg.heads, D, eps))
            else:       # forward_without_attn (eva_vit_model.py:317-324, 249-256): v-projection only, explicit inner LN
                ops.gemm(ws.xb, pb.wqkv_f[2 * D:], ws.att, M=M, bias=pb.c2_qkv[2 * D:], ln_fold=(sx[0], pb.c1_qkv[2 * D:], *sx[1:]))
                ops.layernorm_fwd(ws.att, M, D, pb.gi, pb.bi, eps, u)
                ops.gemm(u, pb.wproj, x, M=M, bias=pb.bproj, residual=x, out2=ws.xb, stats_out=ws.stats_x)
            ops.gemm(ws.xb, pb.w12_f, ws.h, M=M, mode=L.EPI_SWIGLU, bias=pb.c2_w12, stats_out=ws.stats_h,
                     ln_fold=(sx[0], pb.c1_w12, *sx[1:]))
            ops.gemm(ws.h, pb.w3_f, x, M=M, bias=pb.c2_w3, residual=x, out2=ws.xb, stats_out=ws.stats_x,
                     ln_fold=(ws.stats_h, pb.c1_w3, cfg.hidden_pad // 64, cfg.hidden, eps))
            return
        ops.layernorm_fwd(x, M, D, pb.g1, pb.b1, eps, u)
        if with_attention:
            ops.gemm(u, pb.wqkv, ws.qkv, M=M, mode=L.EPI_QKV_ROPE, bias=pb.bqkv, rope=(self.w.rope_pos, self.w.rope_freq),
                     tokens=N, rope_cols=2 * D)
            ops.attention_fwd(ws.qkv, B, N, cfg.heads, self.scale, ws.att, row_stats=ws.stats_att if fold_proj else None)
        else:
            ops.gemm(u, pb.wv, ws.att, M=M, bias=pb.bv)
        if fold_proj:      # inner_attn_ln folded into the proj GEMM's epilogue
            ops.gemm(ws.att, pb.wproj_f, x, M=M, bias=pb.c2_proj, residual=x,
                     ln_fold=(ws.stats_att, pb.c1_proj, 4 * cfg.heads, D, eps))
        else:
            ops.layernorm_fwd(ws.att, M, D, pb.gi, pb.bi, eps, u)
            ops.gemm(u, pb.wproj, x, M=M, bias=pb.bproj, residual=x)
        ops.layernorm_fwd(x, M, D, pb.g2, pb.b2, eps, u)
        if self.fold_w3:   # ffn_ln folded into the w3 GEMM's epilogue
            ops.gemm(u, pb.w12, ws.h, M=M, mode=L.EPI_SWIGLU, bias=pb.b12, stats_out=ws.stats_h)
            ops.gemm(ws.h, pb.w3_f, x, M=M, bias=pb.c2_w3, residual=x,
                     ln_fold=(ws.stats_h, pb.c1_w3, cfg.hidden_pad // 64, cfg.hidden, eps))
        else:
            ops.gemm(u, pb.w12, ws.h, M=M, mode=L.EPI_SWIGLU, bias=pb.b12)
            ops.layernorm_fwd(ws.h, M, cfg.hidden, pb.gf, pb.bf, eps, ws.h2)
            ops.gemm(ws.h2, pb.w3, x, M=M, bias=pb.b3, residual=x)

    # ------------------------------------------------------------------ teacher
    def _cls_chunk(self, images: Tensor, n: int, ws: Workspace, cls_ln: Tensor, out: Tensor) -> None:
        """The kernel sequence of one teacher chunk: n crops -> out [n, embed_dim]."""
        cfg = self.cfg
        self.embed(images, ws.x, ws)
        for i in range(cfg.layers):
            self.block_inplace(i, ws, n)
        ops.layernorm_fwd(ws.x, n, cfg.width, self.w.norm_g, self.w.norm_b, cfg.ln_eps, cls_ln, row_mul=cfg.tokens)
        ops.gemm(cls_ln, self.w.head_w, out, M=n, bias=self.w.head_b)

    def _cls_chunk_graphed(self, images: Tensor, n: int, ws: Workspace, cls_ln: Tensor, out: Tensor) -> None:
        """Same, replayed from a CUDA graph once the (input address, chunk size) pair has been seen twice: the chunk's
        ~65 launches (5 per block) become one graph launch, so small L2-sized chunks cost no host time.  The first call
        of a key runs eagerly (kernel attributes / descriptor cache warm-up), the second captures.  Graphs hold raw
        pointers: they are keyed on the input slice's address and dropped when the workspace or the weights change."""
        if not self.use_graphs or ops.GEMM_PROFILE is not None or torch.cuda.is_current_stream_capturing():
            return self._cls_chunk(images, n, ws, cls_ln, out)
        key = (images.data_ptr(), n, images.dtype, ws.x.data_ptr(), cls_ln.data_ptr())
        ent = self._graphs.get(key)
        if ent is None:
            self._graphs[key] = "seen"
            return self._cls_chunk(images, n, ws, cls_ln, out)
        if ent == "seen":
            if len(self._graphs) > 256:
                self._graphs.clear()
            g = torch.cuda.CUDAGraph()
            g_out = torch.empty(n, self.cfg.embed_dim, device=self.device, dtype=torch.float32)
            l0 = L.launch_count
            torch.cuda.current_stream().synchronize()
            with torch.cuda.graph(g):
                self._cls_chunk(images, n, ws, cls_ln, g_out)
            ent = (g, g_out, L.launch_count - l0, images)        # `images` keeps the captured input storage alive
            self._graphs[key] = ent
        g, g_out, launches, _ = ent
        g.replay()
        L.launch_count += launches
        out.copy_(g_out)

    # ------------------------------------------------------------------ teacher
    def forward_cls(self, images: Tensor, out: Optional[Tensor] = None, ready_events=None) -> Tensor:
        """encode_image(normalize=False): [R,3,S,S] -> [R, embed_dim] f32, no autograd.
        ready_events[k] (optional) gates chunk k on an in-flight H2D copy of its rows."""
        cfg = self.cfg
        R = images.shape[0]
        out = out if out is not None else torch.empty(R, cfg.embed_dim, device=self.device, dtype=torch.float32)
        step = min(self.chunk_images, R)
        ws = self.workspace(step)
        if self._cls_ln is None or self._cls_ln.shape[0] < step:
            self._cls_ln = torch.empty(step, cfg.width, device=self.device, dtype=torch.bfloat16)
        cls_ln = self._cls_ln
        pieces = chunk_schedule(R, step) if ready_events is not None else [(s, min(step, R - s)) for s in range(0, R, step)]
        if ready_events is not None:
            assert len(ready_events) == len(pieces)
        native = self.native is not None and ops.GEMM_PROFILE is None and cfg.grid == self.native.cfg.grid
        for k, (s, n) in enumerate(pieces):
            if ready_events is not None:
                torch.cuda.current_stream().wait_event(ready_events[k])
            if native:
                self.native.forward_cls(images[s:s + n], out[s:s + n], step)
            else:
                self._cls_chunk_graphed(images[s:s + n], n, ws, cls_ln, out[s:s + n])
        return out

    # ------------------------------------------------------------------ student (inference)
    def encode_dense_nograd(self, images: Tensor) -> Tensor:
        """encode_dense: [B,3,S,S] -> NHWC [B,h,w,C] f32, unit-norm per token (no tape)."""
        cfg = self.cfg
        B = images.shape[0]
        g, C = cfg.grid, cfg.embed_dim
        out = torch.empty(B, g, g, C, device=self.device, dtype=torch.float32)
        step = min(self.chunk_images, B)
        if self.native is not None and ops.GEMM_PROFILE is None:
            self.native.forward_dense(images, out, step, cfg.image_size, None if g == self.native.cfg.grid else self.w.pos)
            return out
        ws = self.workspace(step)
        for s in range(0, B, step):
            n = min(step, B - s)
            self.embed(images[s:s + n], ws.x, ws)
            for i in range(cfg.layers - 1):
                self.block_inplace(i, ws, n)
            self.block_inplace(cfg.layers - 1, ws, n, with_attention=False)
            Mp = n * g * g
            tok_ln = ws.u[:Mp]
            ops.layernorm_fwd(ws.x, Mp, cfg.width, self.w.norm_g, self.w.norm_b, cfg.ln_eps, tok_ln,
                              row_div=g * g, row_off=1)
            head = torch.empty(Mp, C, device=self.device, dtype=torch.float32)
            ops.gemm(tok_ln, self.w.head_w, head, M=Mp, bias=self.w.head_b)
            y, _ = ops.l2norm_fwd(head)
            out[s:s + n] = y.view(n, g, g, C)
        return out
