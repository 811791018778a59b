"""Student (trainable) tower: flat parameter / gradient buffers, taped forward, hand-written backward.

Reference semantics being reproduced (wusize/CLIPSelf @ 1c7fe9c):
  forward   EVAVisionTransformer.encode_dense            eva_vit_model.py:588-623
  backward  torch autograd through the above, for the parameters left trainable by
            EVAVisionTransformer.lock(unlocked_groups=n)            eva_vit_model.py:500-516
            (visual.blocks[-n:], n = depth in the scripts; the last block's q_proj/k_proj/q_bias never
            receive a gradient because forward_without_attn skips them, eva_vit_model.py:249-256, 317-324)
  grad sync one mean all-reduce over the trainable range of the flat gradient buffer (what DDP at
            main.py:188 is meant to do) — issued at the end of the backward
            (model.py: _RoiFeatures.backward) on `flat_grad[decay_start(first_trainable):n_grad]`.
  input     any square resolution up to a 64 x 64 token grid (--det-image-size / --multiscale):
            RoPE tables and pos_embed per resolution, rope.py:179-214, eva_vit_model.py:631-643.

Memory layout (HBM, all f32, one allocation each for params / grads / Adam moments):
  [ 2-D weights of every block, in GEMM-friendly groups | 1-D vectors | grad-less tail ]
   '--- weight-decay group (main.py:199-213) ---------'  '- no-decay -'  '- never updated -'
  q|k|v weights of a block are adjacent ([3D,D] view = fused QKV operand and its wgrad output),
  w1|w2 likewise ([2Hd,D]).  The wgrad GEMMs and bias/LN column reductions write straight into
  the flat gradient buffer: no per-tensor copies, one all-reduce, two AdamW launches.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib as L
from . import ops
from .tower import PackedTower, TowerCfg, _round_up, input_grid

Tensor = torch.Tensor

W_NAMES = ["attn.q_proj.weight", "attn.k_proj.weight", "attn.v_proj.weight", "attn.proj.weight",
           "mlp.w1.weight", "mlp.w2.weight", "mlp.w3.weight"]
V_NAMES = ["norm1.weight", "norm1.bias", "attn.q_bias", "attn.v_bias", "attn.inner_attn_ln.weight",
           "attn.inner_attn_ln.bias", "attn.proj.bias", "norm2.weight", "norm2.bias", "mlp.w1.bias",
           "mlp.w2.bias", "mlp.ffn_ln.weight", "mlp.ffn_ln.bias", "mlp.w3.bias"]
GRADLESS_LAST = ("attn.q_proj.weight", "attn.k_proj.weight", "attn.q_bias")


def block_param_shape(cfg: TowerCfg, name: str) -> Tuple[int, ...]:
    D, Hd = cfg.width, cfg.hidden
    if name in ("mlp.w1.weight", "mlp.w2.weight"):
        return (Hd, D)
    if name == "mlp.w3.weight":
        return (D, Hd)
    if name.endswith("proj.weight"):
        return (D, D)
    if name in ("mlp.w1.bias", "mlp.w2.bias", "mlp.ffn_ln.weight", "mlp.ffn_ln.bias"):
        return (Hd,)
    return (D,)


class FlatLayout:
    """Offsets (in elements) of every `blocks.*` parameter inside the flat buffer."""

    def __init__(self, cfg: TowerCfg):
        self.cfg = cfg
        self.offset: Dict[str, int] = {}
        self.shape: Dict[str, Tuple[int, ...]] = {}
        last = cfg.layers - 1
        off = 0
        tail: List[str] = []

        def place(full: str, shape):
            nonlocal off
            off = (off + 3) // 4 * 4                       # every view starts 16-byte aligned (ViT-L: 2730-wide vectors)
            self.offset[full] = off
            self.shape[full] = shape
            n = 1
            for s in shape:
                n *= s
            off += n

        for i in range(cfg.layers):
            for nm in W_NAMES:
                full = f"blocks.{i}.{nm}"
                if i == last and nm in GRADLESS_LAST:
                    tail.append(full)
                else:
                    place(full, block_param_shape(cfg, nm))
        off = (off + 3) // 4 * 4
        self.n_decay = off                      # [0, n_decay): 2-D weights, weight-decay group
        for i in range(cfg.layers):
            for nm in V_NAMES:
                full = f"blocks.{i}.{nm}"
                if i == last and nm in GRADLESS_LAST:
                    tail.append(full)
                else:
                    place(full, block_param_shape(cfg, nm))
        off = (off + 3) // 4 * 4
        self.n_grad = off                       # [n_decay, n_grad): vectors, no-decay group
        for full in tail:
            nm = full.split(".", 2)[2]
            place(full, block_param_shape(cfg, nm))
        self.n_total = off                      # [n_grad, n_total): parameters that never get a gradient
        self.gradless = tuple(tail)

    def decay_start(self, first_block: int) -> int:
        """Start of the weight-decay range of blocks[first_block:] (blocks are laid out in order)."""
        if first_block <= 0:
            return 0
        return min(o for o in (self.offset[f"blocks.{first_block}.{nm}"] for nm in W_NAMES) if o < self.n_decay)

    def nodecay_start(self, first_block: int) -> int:
        if first_block <= 0:
            return self.n_decay
        return min(o for o in (self.offset[f"blocks.{first_block}.{nm}"] for nm in V_NAMES) if o < self.n_grad)

    def view(self, flat: Tensor, name: str) -> Tensor:
        o = self.offset[name]
        shape = self.shape[name]
        n = 1
        for s in shape:
            n *= s
        return flat[o:o + n].view(shape)

    def names(self) -> List[str]:
        return list(self.offset.keys())


def allreduce_flat_gradient(flat_grad: Tensor, layout: FlatLayout, first_trainable: int = 0) -> None:
    """The ONE collective of the step: mean all-reduce of the trainable range of the flat
    gradient buffer — NCCL over NVLink / NVSwitch on the GPUs (`ReduceOp.AVG`); backends without AVG (gloo,
    used by the CPU tests) sum and scale.  The grad-less tail and the weights of frozen blocks are never sent
    (the few 1-D vectors of frozen blocks lie inside the single contiguous span; they are zeros and unused)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    span = flat_grad[layout.decay_start(first_trainable):layout.n_grad]
    if dist.get_backend() == "nccl":
        dist.all_reduce(span, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(span, op=dist.ReduceOp.SUM)
        span /= dist.get_world_size()


class _BlockPack:
    __slots__ = ("wqkv", "wqkvT", "bqkv", "wv", "wvT", "wproj", "wprojT", "w12", "w12T", "b12", "w3", "w3T")


class _Tape:
    """Saved activations of one student forward (allocated once per batch size)."""

    def __init__(self, cfg: TowerCfg, B: int, dev, grid: Optional[int] = None):
        grid = grid or cfg.grid
        D, Hd, N, H, Lr = cfg.width, cfg.hidden_pad, grid * grid + 1, cfg.heads, cfg.layers   # Hd: padded hidden width
        self.grid, self.N = grid, N
        M = B * N
        Mp = B * (N - 1)
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        self.B, self.M, self.Mp = B, M, Mp
        self.x = [torch.empty(M, D, **f32) for _ in range(Lr + 1)]
        self.xmid = [torch.empty(M, D, **f32) for _ in range(Lr)]
        self.u = [torch.empty(M, D, **bf) for _ in range(Lr)]
        self.qkv = [torch.empty(M, 3 * D, **bf) if i < Lr - 1 else None for i in range(Lr)]
        self.att = [torch.empty(M, D, **bf) for _ in range(Lr)]
        self.lse = [torch.empty(B, H, N, **f32) if i < Lr - 1 else None for i in range(Lr)]
        self.aln = [torch.empty(M, D, **bf) for _ in range(Lr)]
        self.u2 = [torch.empty(M, D, **bf) for _ in range(Lr)]
        self.x12 = [torch.empty(M, 2 * Hd, **bf) for _ in range(Lr)]
        # padded hidden columns are never written by the LayerNorm / SwiGLU kernels: keep them exact zeros
        self.h = [torch.zeros(M, Hd, **bf) for _ in range(Lr)]
        self.hln = [torch.zeros(M, Hd, **bf) for _ in range(Lr)]
        self.stats = [[torch.empty(M, **f32) for _ in range(8)] for _ in range(Lr)]   # mean/rstd x 4 LNs
        self.tok_ln = torch.empty(Mp, D, **bf)
        self.tok_stats = [torch.empty(Mp, **f32) for _ in range(2)]
        self.head = torch.empty(Mp, cfg.embed_dim, **f32)
        self.dense = torch.empty(Mp, cfg.embed_dim, **f32)
        self.inv_norm = torch.empty(Mp, **f32)
        # backward scratch
        Mpad = _round_up(M, 8)
        self.Mpad = Mpad
        self.dx = torch.empty(M, D, **f32)
        self.g_bf_D = torch.empty(M, D, **bf)            # bf16 copy of a [M,D] gradient
        self.g_Hd = torch.empty(M, Hd, **bf)
        self.g_Hd2 = torch.empty(M, Hd, **bf)
        self.g_2Hd = torch.zeros(M, 2 * Hd, **bf)
        self.g_3D = torch.empty(M, 3 * D, **bf)
        self.g_D2 = torch.empty(M, D, **bf)
        self.delta = torch.empty(B * H * N, **f32)
        self.d_head = torch.empty(Mp, cfg.embed_dim, **f32)
        self.d_head_bf = torch.empty(Mp, cfg.embed_dim, **bf)
        self.col_ws = torch.empty(128 * 2 * max(3 * D, 2 * Hd), **f32)
        self.dw3_pad = torch.zeros(D, Hd, **f32) if Hd != cfg.hidden else None    # wgrad scratch with aligned rows


class StudentEngine:
    """Owns the flat f32 parameter / gradient buffers of `visual.blocks.*` and runs the taped
    forward and the backward of the dense path."""

    def __init__(self, cfg: TowerCfg, sd: Dict[str, Tensor], device: torch.device):
        L.require_device()
        assert cfg.head_dim == 64 and cfg.hidden % 2 == 0
        self.Hd, self.Hp = cfg.hidden, cfg.hidden_pad           # real / padded SwiGLU hidden width
        self.padded = self.Hp != self.Hd
        self.cfg, self.device = cfg, device
        self.layout = FlatLayout(cfg)
        lay = self.layout
        self.flat_param = torch.empty(lay.n_total, device=device, dtype=torch.float32)
        self.flat_grad = torch.zeros(lay.n_total, device=device, dtype=torch.float32)
        for name in lay.names():
            lay.view(self.flat_param, name).copy_(sd[name].detach().to(device=device, dtype=torch.float32))
        # frozen parts (patch embed, cls, pos, final norm, head): the generic packer handles them
        self.frozen = PackedTower.__new__(PackedTower)
        self._pack_frozen(sd)
        self.packs = [_BlockPack() for _ in range(cfg.layers)]
        self._alloc_packs()
        self.scale = cfg.head_dim ** -0.5
        self._tape: Optional[_Tape] = None
        # blocks[:first_trainable] are frozen (lock_image_tower(unlocked_groups=n) unfreezes blocks[-n:],
        # eva_vit_model.py:500-516): no gradient, no all-reduce, no optimizer update for them
        self.first_trainable = 0
        self.weights_epoch = 0          # bumped by FusedAdamW.step(): in-place updates through the C ABI are invisible to torch
        # overlapped gradient exchange (opt-in, see allreduce_async): the all-reduce and the fused AdamW run on this side
        # stream while the main stream already runs the NEXT step's frozen teacher
        self.comm_stream: Optional[torch.cuda.Stream] = None
        self.grads_ready: Optional[torch.cuda.Event] = None      # recorded on comm_stream after the all-reduce
        self.weights_ready: Optional[torch.cuda.Event] = None    # recorded on comm_stream after the fused AdamW
        self.repack()

    # ------------------------------------------------------------------ packing
    def _pack_frozen(self, sd):
        cfg, dev = self.cfg, self.device
        f = lambda k: sd[k].detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        fr = self.frozen
        from .tower import rope_tables, rope_vectors
        self._res: Dict[int, SimpleNamespace] = {}
        fr.pos_src = f("pos_embed")
        fr.k_pe = 3 * cfg.patch * cfg.patch
        fr.k_pe_pad = _round_up(fr.k_pe, 8)
        D = cfg.width
        fr.pe_w = ops.cast_pad_bf16(f("patch_embed.proj.weight").reshape(D, -1), fr.k_pe_pad)
        fr.pe_b = f("patch_embed.proj.bias")
        fr.cls = f("cls_token").reshape(-1)
        fr.norm_g, fr.norm_b = f("norm.weight"), f("norm.bias")
        hw = f("head.weight")
        fr.head_w = torch.empty(cfg.embed_dim, D, device=dev, dtype=torch.bfloat16)
        self.head_wT = torch.empty(D, cfg.embed_dim, device=dev, dtype=torch.bfloat16)
        ops.cast_transpose(hw, cfg.embed_dim, D, dst=fr.head_w, dst_t=self.head_wT)
        fr.head_b = f("head.bias")

    def _alloc_packs(self):
        cfg, dev = self.cfg, self.device
        D, Hd = cfg.width, cfg.hidden_pad
        bf = dict(device=dev, dtype=torch.bfloat16)
        for i, pk in enumerate(self.packs):
            last = i == cfg.layers - 1
            if not last:
                pk.wqkv = torch.empty(3 * D, D, **bf)
                pk.wqkvT = torch.empty(D, 3 * D, **bf)
                pk.bqkv = torch.zeros(3 * D, device=dev, dtype=torch.float32)
                pk.wv = pk.wvT = None
            else:
                pk.wqkv = pk.wqkvT = pk.bqkv = None
                pk.wv = torch.empty(D, D, **bf)
                pk.wvT = torch.empty(D, D, **bf)
            pk.wproj = torch.empty(D, D, **bf)
            pk.wprojT = torch.empty(D, D, **bf)
            # zero-initialised: the padded rows / columns (ViT-L) are never rewritten
            pk.w12 = torch.zeros(2 * Hd, D, **bf)
            pk.w12T = torch.zeros(D, 2 * Hd, **bf)
            pk.b12 = torch.zeros(2 * Hd, device=dev, dtype=torch.float32)
            pk.w3 = torch.zeros(D, Hd, **bf)
            pk.w3T = torch.zeros(Hd, D, **bf)

    def p(self, i: int, name: str) -> Tensor:
        return self.layout.view(self.flat_param, f"blocks.{i}.{name}")

    def g(self, i: int, name: str) -> Tensor:
        return self.layout.view(self.flat_grad, f"blocks.{i}.{name}")

    def _span(self, flat: Tensor, i: int, first: str, rows: int, cols: int) -> Tensor:
        o = self.layout.offset[f"blocks.{i}.{first}"]
        return flat[o:o + rows * cols].view(rows, cols)

    # ------------------------------------------------------------------ overlapped gradient exchange
    def allreduce_async(self) -> None:
        """The step's ONE collective, issued on a side stream right after the last wgrad: the mean all-reduce (NCCL over
        NVLink / NVSwitch) of the flat gradient no longer sits at the end of the main stream's critical path.  Consumers
        order themselves after `grads_ready` (FusedAdamW.step runs on the same side stream; anything that reads the
        gradients on another stream calls wait_gradients())."""
        cur = torch.cuda.current_stream()
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=self.device)
        self.comm_stream.wait_stream(cur)                       # the backward has written every gradient
        with torch.cuda.stream(self.comm_stream):
            allreduce_flat_gradient(self.flat_grad, self.layout, self.first_trainable)
            self.grads_ready = self.comm_stream.record_event()

    def wait_gradients(self) -> None:
        if self.grads_ready is not None:
            torch.cuda.current_stream().wait_event(self.grads_ready)

    def wait_weights(self) -> None:
        """Order the current stream after a fused optimizer step that ran on the side stream."""
        if self.weights_ready is not None:
            torch.cuda.current_stream().wait_event(self.weights_ready)
            self.weights_ready = None
            self.grads_ready = None

    def repack(self) -> None:
        """f32 master -> bf16 GEMM operands (W and W^T), after every optimizer step."""
        self.wait_weights()
        cfg = self.cfg
        D, Hd = cfg.width, cfg.hidden
        for i, pk in enumerate(self.packs):
            last = i == cfg.layers - 1
            if not last:
                ops.cast_transpose(self._span(self.flat_param, i, "attn.q_proj.weight", 3 * D, D), 3 * D, D,
                                   dst=pk.wqkv, dst_t=pk.wqkvT)
                pk.bqkv[:D].copy_(self.p(i, "attn.q_bias"))
                pk.bqkv[2 * D:].copy_(self.p(i, "attn.v_bias"))
            else:
                ops.cast_transpose(self.p(i, "attn.v_proj.weight"), D, D, dst=pk.wv, dst_t=pk.wvT)
            ops.cast_transpose(self.p(i, "attn.proj.weight"), D, D, dst=pk.wproj, dst_t=pk.wprojT)
            Hp = self.Hp
            if not self.padded:
                ops.cast_transpose(self._span(self.flat_param, i, "mlp.w1.weight", 2 * Hd, D), 2 * Hd, D,
                                   dst=pk.w12, dst_t=pk.w12T)
            else:   # [w1 | pad | w2 | pad] rows, transposed copy column blocks at 0 and Hp
                ops.cast_transpose(self.p(i, "mlp.w1.weight"), Hd, D, dst=pk.w12[:Hd], dst_t=pk.w12T, ldt=2 * Hp)
                ops.cast_transpose(self.p(i, "mlp.w2.weight"), Hd, D, dst=pk.w12[Hp:Hp + Hd], dst_t=pk.w12T[:, Hp:], ldt=2 * Hp)
            pk.b12[:Hd].copy_(self.p(i, "mlp.w1.bias"))
            pk.b12[Hp:Hp + Hd].copy_(self.p(i, "mlp.w2.bias"))
            ops.cast_transpose(self.p(i, "mlp.w3.weight"), D, Hd, dst=pk.w3, dst_t=pk.w3T)

    # ------------------------------------------------------------------ forward
    def resolution(self, grid: int) -> SimpleNamespace:
        """Per-resolution constants: RoPE tables / vectors with ft_seq_len = grid (rope.py:179-214) and the
        bicubically rescaled pos_embed (eva_vit_model.py:631-643); the student runs at --det-image-size."""
        r = self._res.get(grid)
        if r is None:
            from .tower import rescale_pos_embed, rope_tables, rope_vectors
            cfg, dev = self.cfg, self.device
            r = SimpleNamespace()
            cos, sin = rope_tables(grid, cfg.head_dim, cfg.pt_seq_len)
            r.rope_cos, r.rope_sin = cos.to(dev), sin.to(dev)
            pos, freq = rope_vectors(grid, cfg.head_dim, cfg.pt_seq_len)
            r.rope_pos, r.rope_freq = pos.to(dev), freq.to(dev)
            r.pos = rescale_pos_embed(self.frozen.pos_src, grid)
            self._res[grid] = r
        return r

    def tape(self, B: int, grid: Optional[int] = None) -> _Tape:
        grid = grid or self.cfg.grid
        if self._tape is None or self._tape.B != B or self._tape.grid != grid:
            self._tape = None                       # release the old activations before allocating the new ones
            self._tape = _Tape(self.cfg, B, self.device, grid)
        return self._tape

    def forward(self, images: Tensor) -> Tensor:
        """encode_dense with everything the backward needs kept on the tape.
        Returns the NHWC map [B,g,g,C] f32 (a view of the tape)."""
        cfg, fr = self.cfg, self.frozen
        g = input_grid(images, cfg)
        res = self.resolution(g)
        D, Hd, N = cfg.width, cfg.hidden, g * g + 1
        B = images.shape[0]
        t = self.tape(B, g)
        M = t.M
        eps = cfg.ln_eps
        # embed (frozen): same kernels as the teacher
        patches = ops.im2col_patches(images, cfg.patch, fr.k_pe_pad)
        ops.gemm(patches, fr.pe_w, t.x[0], M=B * (N - 1), N=D, K=fr.k_pe_pad, mode=L.EPI_TOKENS, bias=fr.pe_b,
                 pos_embed=res.pos, tokens=N)
        ops.fill_cls_rows(fr.cls, res.pos, t.x[0].view(B, N, D))
        for i, pk in enumerate(self.packs):
            last = i == cfg.layers - 1
            st = t.stats[i]
            x = t.x[i]
            ops.layernorm_fwd(x, M, D, self.p(i, "norm1.weight"), self.p(i, "norm1.bias"), eps, t.u[i], mean=st[0], rstd=st[1])
            if not last:
                ops.gemm(t.u[i], pk.wqkv, t.qkv[i], M=M, mode=L.EPI_QKV_ROPE, bias=pk.bqkv,
                         rope=(res.rope_pos, res.rope_freq), tokens=N, rope_cols=2 * D)
                ops.attention_fwd(t.qkv[i], B, N, cfg.heads, self.scale, t.att[i], t.lse[i])
            else:
                ops.gemm(t.u[i], pk.wv, t.att[i], M=M, bias=self.p(i, "attn.v_bias"))
            ops.layernorm_fwd(t.att[i], M, D, self.p(i, "attn.inner_attn_ln.weight"), self.p(i, "attn.inner_attn_ln.bias"),
                              eps, t.aln[i], mean=st[2], rstd=st[3])
            ops.gemm(t.aln[i], pk.wproj, t.xmid[i], M=M, bias=self.p(i, "attn.proj.bias"), residual=x)
            ops.layernorm_fwd(t.xmid[i], M, D, self.p(i, "norm2.weight"), self.p(i, "norm2.bias"), eps, t.u2[i],
                              mean=st[4], rstd=st[5])
            ops.gemm(t.u2[i], pk.w12, t.x12[i], M=M, bias=pk.b12)
            ops.swiglu_fwd(t.x12[i], M, Hd, t.h[i], split=self.Hp if self.padded else True)
            ops.layernorm_fwd(t.h[i], M, Hd, self.p(i, "mlp.ffn_ln.weight"), self.p(i, "mlp.ffn_ln.bias"), eps, t.hln[i],
                              mean=st[6], rstd=st[7])
            ops.gemm(t.hln[i], pk.w3, t.x[i + 1], M=M, bias=self.p(i, "mlp.w3.bias"), residual=t.xmid[i])
        ops.layernorm_fwd(t.x[cfg.layers], t.Mp, D, fr.norm_g, fr.norm_b, eps, t.tok_ln, row_div=g * g, row_off=1,
                          mean=t.tok_stats[0], rstd=t.tok_stats[1])
        ops.gemm(t.tok_ln, fr.head_w, t.head, M=t.Mp, bias=fr.head_b)
        L.call("cs_l2norm_fwd", t.head.data_ptr(), t.Mp, cfg.embed_dim, t.dense.data_ptr(), t.inv_norm.data_ptr(),
               torch.cuda.current_stream().cuda_stream)
        return t.dense.view(B, g, g, cfg.embed_dim)

    # ------------------------------------------------------------------ backward
    def backward(self, d_dense: Tensor) -> None:
        """Gradient of the dense map w.r.t. every trainable block parameter, written into
        self.flat_grad[:n_grad] (overwritten, not accumulated: accum_freq == 1, train.py:89)."""
        cfg, fr, t = self.cfg, self.frozen, self._tape
        D, Hd, N, H = cfg.width, cfg.hidden, t.N, cfg.heads
        B, M, Mp = t.B, t.M, t.Mp
        g2 = t.grid * t.grid
        res = self.resolution(t.grid)
        ws = t.col_ws
        stream = torch.cuda.current_stream().cuda_stream
        # tail: normalise -> head -> final LN (all frozen: input gradients only)
        L.call("cs_l2norm_bwd", t.dense.data_ptr(), t.inv_norm.data_ptr(), d_dense.data_ptr(), Mp, cfg.embed_dim,
               t.d_head.data_ptr(), stream)
        ops.cast_transpose(t.d_head, Mp, cfg.embed_dim, dst=t.d_head_bf)
        d_tok = t.g_D2[:Mp]
        ops.gemm(t.d_head_bf, self.head_wT, d_tok, M=Mp)
        k0 = self.first_trainable
        self.flat_grad[self.layout.decay_start(k0):self.layout.n_decay].zero_()   # split-K wgrad GEMMs accumulate with red.add
        dx = t.dx
        dx.zero_()                                                     # CLS rows get no gradient from the tail
        ops.layernorm_bwd_dx(d_tok, t.x[cfg.layers], Mp, D, t.tok_stats[0], t.tok_stats[1], fr.norm_g, dx,
                             row_div=g2, row_off=1, row_mapped=True)
        for i in range(cfg.layers - 1, k0 - 1, -1):
            pk, st = self.packs[i], t.stats[i]
            last = i == cfg.layers - 1
            # ---- MLP branch: x_out = xmid + w3(hln) + b3
            # weight gradients: dW = dY^T X on dY / X as they lie in memory (cs_gemm_bf16_tn: both operands MN-major)
            ops.cast_transpose(dx, M, D, dst=t.g_bf_D)
            ops.col_reduce(dx, M, D, self.g(i, "mlp.w3.bias"), ws)
            if not self.padded:
                ops.gemm_tn(t.g_bf_D, t.hln[i], self.g(i, "mlp.w3.weight"), M=D, N=Hd, K=M, k_splits=-1)    # dW3 = dx^T hln
            else:   # rows of the [D, 2730] gradient are not 16 B aligned: GEMM into a padded scratch, then copy
                t.dw3_pad.zero_()
                ops.gemm_tn(t.g_bf_D, t.hln[i], t.dw3_pad, M=D, N=self.Hp, K=M, k_splits=-1)
                self.g(i, "mlp.w3.weight").copy_(t.dw3_pad[:, :Hd])
            ops.gemm(t.g_bf_D, pk.w3T, t.g_Hd, M=M)                                               # d_hln
            ops.col_reduce(t.g_Hd, M, Hd, self.g(i, "mlp.ffn_ln.bias"), ws, x=t.h[i], mean=st[6], rstd=st[7],
                           dgamma=self.g(i, "mlp.ffn_ln.weight"))
            ops.layernorm_bwd_dx(t.g_Hd, t.h[i], M, Hd, st[6], st[7], self.p(i, "mlp.ffn_ln.weight"), t.g_Hd2)  # d_h
            ops.swiglu_bwd(t.x12[i], t.g_Hd2, M, Hd, t.g_2Hd, split=self.Hp if self.padded else True)   # d_x12
            if not self.padded:
                ops.gemm_tn(t.g_2Hd, t.u2[i], self._span(self.flat_grad, i, "mlp.w1.weight", 2 * Hd, D),
                            M=2 * Hd, N=D, K=M, k_splits=-1)                                      # dW1|dW2
                ops.col_reduce(t.g_2Hd, M, 2 * Hd, self._span(self.flat_grad, i, "mlp.w1.bias", 1, 2 * Hd).view(-1), ws)
            else:
                Hp = self.Hp
                ops.gemm_tn(t.g_2Hd, t.u2[i], self.g(i, "mlp.w1.weight"), M=Hd, N=D, K=M, lda=2 * Hp, k_splits=-1)
                ops.gemm_tn(t.g_2Hd[:, Hp:], t.u2[i], self.g(i, "mlp.w2.weight"), M=Hd, N=D, K=M, lda=2 * Hp, k_splits=-1)
                ops.col_reduce(t.g_2Hd, M, Hd, self.g(i, "mlp.w1.bias"), ws, lddy=2 * Hp)
                ops.col_reduce(t.g_2Hd[:, Hp:], M, Hd, self.g(i, "mlp.w2.bias"), ws, lddy=2 * Hp)
            ops.gemm(t.g_2Hd, pk.w12T, t.g_D2, M=M)                                               # d_u2
            ops.col_reduce(t.g_D2, M, D, self.g(i, "norm2.bias"), ws, x=t.xmid[i], mean=st[4], rstd=st[5],
                           dgamma=self.g(i, "norm2.weight"))
            ops.layernorm_bwd_dx(t.g_D2, t.xmid[i], M, D, st[4], st[5], self.p(i, "norm2.weight"), dx, add=dx)  # d_xmid
            # ---- attention branch: xmid = x + proj(aln) + b
            ops.cast_transpose(dx, M, D, dst=t.g_bf_D)
            ops.col_reduce(dx, M, D, self.g(i, "attn.proj.bias"), ws)
            ops.gemm_tn(t.g_bf_D, t.aln[i], self.g(i, "attn.proj.weight"), M=D, N=D, K=M, k_splits=-1)  # dWproj
            ops.gemm(t.g_bf_D, pk.wprojT, t.g_D2, M=M)                                            # d_aln
            ops.col_reduce(t.g_D2, M, D, self.g(i, "attn.inner_attn_ln.bias"), ws, x=t.att[i], mean=st[2], rstd=st[3],
                           dgamma=self.g(i, "attn.inner_attn_ln.weight"))
            d_att = t.g_bf_D
            ops.layernorm_bwd_dx(t.g_D2, t.att[i], M, D, st[2], st[3], self.p(i, "attn.inner_attn_ln.weight"), d_att)
            if not last:
                ops.attention_bwd(t.qkv[i], t.att[i], d_att, t.lse[i], B, N, H, self.scale,
                                  (res.rope_cos, res.rope_sin), t.delta, t.g_3D)                    # d_qkv (raw projections)
                ops.gemm_tn(t.g_3D, t.u[i], self._span(self.flat_grad, i, "attn.q_proj.weight", 3 * D, D),
                            M=3 * D, N=D, K=M, k_splits=-1)                                       # dWq|dWk|dWv
                ops.col_reduce(t.g_3D, M, D, self.g(i, "attn.q_bias"), ws, lddy=3 * D)
                ops.col_reduce(t.g_3D[:, 2 * D:], M, D, self.g(i, "attn.v_bias"), ws, lddy=3 * D)
                ops.gemm(t.g_3D, pk.wqkvT, t.g_D2, M=M)                                           # d_u
            else:
                ops.gemm_tn(d_att, t.u[i], self.g(i, "attn.v_proj.weight"), M=D, N=D, K=M, k_splits=-1)
                ops.col_reduce(d_att, M, D, self.g(i, "attn.v_bias"), ws)
                ops.gemm(d_att, pk.wvT, t.g_D2, M=M)
            ops.col_reduce(t.g_D2, M, D, self.g(i, "norm1.bias"), ws, x=t.x[i], mean=st[0], rstd=st[1],
                           dgamma=self.g(i, "norm1.weight"))
            if i > k0:  # everything below the first trainable block is frozen: its input gradient is not needed
                ops.layernorm_bwd_dx(t.g_D2, t.x[i], M, D, st[0], st[1], self.p(i, "norm1.weight"), dx, add=dx)
