"""Model-level API of the reference, re-hosted on the B200 kernels.

Mirrors (wusize/CLIPSelf @ 1c7fe9c):
  EVAVisionTransformer  src/open_clip/eva_clip/eva_vit_model.py:396-711  (parameter names / shapes,
                        init :455-495, lock :500-516, forward :581-586, encode_dense :588-623,
                        extract_roi_features :625-629, mask_pool :645-653)
  CustomCLIP            src/open_clip/eva_clip/model.py:272-346 (encode_image / encode_dense /
                        encode_pseudo_boxes / encode_masks / lock_image_tower / logit_scale)
The modules below only HOLD parameters under the reference's state_dict keys; every arithmetic
step of the vision tower is executed by the CUDA library through TowerEngine / StudentEngine.
There is no PyTorch fallback: calling these methods without a B200 raises.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .student import StudentEngine, allreduce_flat_gradient
from .tower import TowerCfg, TowerEngine, input_grid, rope_tables

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# Parameter containers (state_dict-compatible with the reference)
# ----------------------------------------------------------------------------------------------
class _Rope(nn.Module):
    """Holds the freqs_cos / freqs_sin buffers of VisionRotaryEmbeddingFast (rope.py:96-146)."""

    def __init__(self, grid: int, head_dim: int, pt_seq_len: int):
        super().__init__()
        cos, sin = rope_tables(grid, head_dim, pt_seq_len)
        self.register_buffer("freqs_cos", cos)
        self.register_buffer("freqs_sin", sin)


class _Attention(nn.Module):
    def __init__(self, dim: int, rope: _Rope):
        super().__init__()
        self.q_proj = nn.Linear(dim, dim, bias=False)
        self.k_proj = nn.Linear(dim, dim, bias=False)
        self.v_proj = nn.Linear(dim, dim, bias=False)
        self.q_bias = nn.Parameter(torch.zeros(dim))
        self.v_bias = nn.Parameter(torch.zeros(dim))
        self.inner_attn_ln = nn.LayerNorm(dim, eps=1e-6)
        self.proj = nn.Linear(dim, dim)
        self.rope = rope


class _SwiGLU(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.w1 = nn.Linear(dim, hidden)
        self.w2 = nn.Linear(dim, hidden)
        self.ffn_ln = nn.LayerNorm(hidden, eps=1e-6)
        self.w3 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim: int, hidden: int, rope: _Rope):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, rope)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _SwiGLU(dim, hidden)


class _PatchEmbed(nn.Module):
    def __init__(self, img_size: int, patch_size: int, embed_dim: int):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.patch_shape = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.patch_shape[0] * self.patch_shape[1]
        self.proj = nn.Conv2d(3, embed_dim, kernel_size=patch_size, stride=patch_size)


def _trunc_normal_(t: Tensor, std: float) -> None:
    nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


# ----------------------------------------------------------------------------------------------
# autograd glue: one node for the whole student dense+RoI path
# ----------------------------------------------------------------------------------------------
def _flat_grad_outputs(visual, eng: StudentEngine) -> list:
    """Publish the gradients of one backward: `.grad` of every trainable blocks.* parameter is pointed at its
    view of the flat gradient buffer the backward has just overwritten, and autograd gets None for them.
    Handing the views to AccumulateGrad instead would add a view to itself (doubling the gradient) whenever
    `.grad` survives from the previous step (optimizer.zero_grad(set_to_none=False), or no zero_grad at all).
    `.grad` therefore always holds the gradient of the latest backward (accum_freq == 1, train.py:89)."""
    grads = []
    for name, p in visual._block_params():
        grads.append(None)
        if name in eng.layout.gradless or not p.requires_grad:
            continue
        v = eng.layout.view(eng.flat_grad, name)
        if p.grad is None or p.grad.data_ptr() != v.data_ptr():
            p.grad = v
    return grads


class _RoiFeatures(torch.autograd.Function):
    """encode_dense -> RoIAlign.  forward/backward run entirely in the CUDA library; the node
    exposes the flat-buffer gradient views to autograd so optimizers see ordinary `.grad`s."""

    @staticmethod
    def forward(ctx, visual, images, rois, img_offsets, R, *params):
        eng: StudentEngine = visual._student_engine()
        dense = eng.forward(images)
        out, wy, wx = ops.roi_align_fwd(dense, rois, img_offsets, R)
        ctx.visual, ctx.R = visual, R
        ctx.shape = tuple(dense.shape)
        ctx.save_for_backward(img_offsets, wy, wx)
        return out

    @staticmethod
    def backward(ctx, d_out):
        visual = ctx.visual
        eng: StudentEngine = visual._student
        img_offsets, wy, wx = ctx.saved_tensors
        d_dense = ops.roi_align_bwd(d_out.contiguous(), ctx.shape, img_offsets, ctx.R, wy, wx)
        eng.backward(d_dense)
        if visual.sync_gradients:
            # the ONE collective of the step: mean all-reduce of the flat student gradient (NCCL/NVLink); with
            # overlap_gradient_sync on a side stream, so that the next step's teacher forward hides it
            if visual.overlap_gradient_sync:
                eng.allreduce_async()
            else:
                allreduce_flat_gradient(eng.flat_grad, eng.layout, eng.first_trainable)
        grads = _flat_grad_outputs(visual, eng)
        return (None, None, None, None, None, *grads)


class _DenseFeatures(torch.autograd.Function):
    """encode_dense with gradient (for callers that pool the map themselves)."""

    @staticmethod
    def forward(ctx, visual, images, *params):
        eng: StudentEngine = visual._student_engine()
        ctx.visual = visual
        return eng.forward(images).clone()

    @staticmethod
    def backward(ctx, d_dense):
        visual = ctx.visual
        eng: StudentEngine = visual._student
        eng.backward(d_dense.contiguous())
        grads = _flat_grad_outputs(visual, eng)
        return (None, None, *grads)


class _CosineLoss(torch.autograd.Function):
    """(1 - mean cos(student, teacher)) * weight  — clipself.py:42-47 fused."""

    @staticmethod
    def forward(ctx, student, teacher, weight):
        student, teacher = student.contiguous(), teacher.contiguous()
        loss, stats = ops.cosine_loss_fwd(student, teacher, weight)
        ctx.save_for_backward(student, teacher, stats)
        ctx.weight = weight
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        student, teacher, stats = ctx.saved_tensors
        d_s = ops.cosine_loss_bwd(student, teacher, stats, ctx.weight, d_loss.contiguous().float())
        return d_s, None, None


def cosine_distill_loss(student: Tensor, teacher: Tensor, weight: float = 1.0) -> Tensor:
    return _CosineLoss.apply(student, teacher, float(weight))


# ----------------------------------------------------------------------------------------------
# EVAVisionTransformer
# ----------------------------------------------------------------------------------------------
class EVAVisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, num_classes=512, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=2.6667, pt_hw_seq_len=16, init_scale=0.001, **unused):
        super().__init__()
        self.image_size = img_size
        self.num_heads = num_heads
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        hidden = int(embed_dim * mlp_ratio)
        self.cfg = TowerCfg(image_size=img_size, patch=patch_size, width=embed_dim, heads=num_heads, layers=depth,
                            hidden=hidden, embed_dim=num_classes, pt_seq_len=pt_hw_seq_len)
        self.patch_embed = _PatchEmbed(img_size, patch_size, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.rope = _Rope(img_size // patch_size, embed_dim // num_heads, pt_hw_seq_len)
        self.blocks = nn.ModuleList([_Block(embed_dim, hidden, self.rope) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.head = nn.Linear(embed_dim, num_classes)
        self.grad_checkpointing = False
        self.sync_gradients = False          # set by the CLIPSelf plug-in when `distributed`
        # opt-in (training CLI / bench, which step with FusedAdamW): gradient all-reduce + optimizer on a side stream,
        # overlapped with the next step's teacher forward.  Off by default: a caller that reads `.grad` right after
        # backward() on its own stream would otherwise have to call visual._student.wait_gradients() first.
        self.overlap_gradient_sync = False
        self._student: Optional[StudentEngine] = None
        self._infer: Optional[TowerEngine] = None
        self._infer_version = None
        self._init_weights(init_scale)

    # -------------------------------------------------------------- construction helpers
    def _init_weights(self, init_scale: float) -> None:
        """trunc-normal(.02) linears, zero biases, unit LN (eva_vit_model.py:455-495), proj/w3
        rescaled by 1/sqrt(2*layer) (:474-484), head scaled by init_scale (:461-467)."""
        _trunc_normal_(self.pos_embed, 0.02)
        _trunc_normal_(self.cls_token, 0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight, 0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        for i, blk in enumerate(self.blocks):
            blk.attn.proj.weight.data.div_(math.sqrt(2.0 * (i + 1)))
            blk.mlp.w3.weight.data.div_(math.sqrt(2.0 * (i + 1)))
        _trunc_normal_(self.head.weight, 0.02)
        self.head.weight.data.mul_(init_scale)
        self.head.bias.data.mul_(init_scale)

    def get_num_layers(self) -> int:
        return len(self.blocks)

    def get_cast_dtype(self) -> torch.dtype:
        return self.blocks[0].mlp.w3.weight.dtype

    def set_grad_checkpointing(self, enable: bool = True) -> None:
        self.grad_checkpointing = enable      # activations are taped explicitly; flag kept for API parity

    def no_weight_decay(self):
        return {"pos_embed", "cls_token"}

    def lock(self, unlocked_groups: int = 0, freeze_bn_stats: bool = False) -> None:
        """Freeze everything, then unfreeze blocks[-unlocked_groups:] (eva_vit_model.py:500-516;
        note unlocked_groups=0 unfreezes ALL blocks, as `blocks[-0:]` does in the reference)."""
        for p in self.parameters():
            p.requires_grad = False
        for blk in self.blocks[-unlocked_groups:]:
            for p in blk.parameters():
                p.requires_grad = True

    # -------------------------------------------------------------- engines
    def _tower_sd(self) -> Dict[str, Tensor]:
        return {k: v for k, v in self.state_dict().items() if "rope" not in k}

    def _block_params(self):
        """(name, parameter) of every blocks.* parameter, in the order used by the autograd nodes."""
        out = []
        for i, blk in enumerate(self.blocks):
            for n, p in blk.named_parameters():
                out.append((f"blocks.{i}.{n}", p))
        return out

    def _device(self) -> torch.device:
        return self.cls_token.device

    def _check_cuda(self):
        if self._device().type != "cuda":
            raise L.ClipselfB200Error("clipself_b200 runs on a B200 only: move the model to cuda (no CPU path)")

    def _student_engine(self) -> StudentEngine:
        """Flat-buffer engine for the trainable tower; block parameters are re-pointed at views of
        the flat f32 buffer so any optimizer updates it in place."""
        self._check_cuda()
        flags = [all(p.requires_grad for p in blk.parameters()) for blk in self.blocks]
        mixed = [any(p.requires_grad for p in blk.parameters()) for blk in self.blocks]
        first = flags.index(True) if True in flags else len(flags)
        if first == len(flags) or flags != mixed or not all(flags[first:]):
            raise NotImplementedError("the fused training path needs a trainable suffix of whole blocks "
                                      "(lock_image_tower(unlocked_groups=n) unfreezes blocks[-n:], as the reference does)")
        stray = [n for n, p in self.named_parameters() if p.requires_grad and not n.startswith("blocks.")]
        if stray:
            # the reference would train these (main.py:199-213 puts every requires_grad parameter in the optimizer); this
            # path treats everything outside the blocks as frozen, so refuse instead of silently not updating them
            raise NotImplementedError(f"trainable parameters outside the transformer blocks are not supported: {stray[:4]} ...; "
                                      "call lock_image_tower(unlocked_groups=n) (every CLIPSelf script passes --lock-image)")
        if self._student is None:
            self._student = StudentEngine(self.cfg, self._tower_sd(), self._device())
        eng = self._student
        for name, p in self._block_params():
            v = eng.layout.view(eng.flat_param, name)
            if p.data_ptr() != v.data_ptr():
                v.copy_(p.data)
                p.data = v
        eng.first_trainable = first
        eng.repack()
        return eng

    def _infer_engine(self, x: Optional[Tensor] = None) -> TowerEngine:
        """Forward-only engine; with `x` given, the view of it for x's resolution (any square multiple of
        the patch size, like the reference tower: eva_vit_model.py:533-549, 631-643, rope.py:179-214)."""
        eng = self._infer_engine_native()
        return eng if x is None else eng.at_grid(input_grid(x, self.cfg))

    def _infer_engine_native(self) -> TowerEngine:
        self._check_cuda()
        # weights_epoch: the fused optimizer updates the flat parameter buffer through the C ABI, which neither bumps
        # `_version` nor moves `data_ptr` — FusedAdamW.step() counts its updates on the engine instead
        version = tuple(p._version for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters()) + \
            (self._student.weights_epoch if self._student is not None else 0,)
        if self._infer is None or version != self._infer_version:
            if self._infer is None:
                self._infer = TowerEngine(self.cfg, self._tower_sd(), self._device())
            else:
                self._infer.repack(self._tower_sd())
            self._infer_version = version
        return self._infer

    def mark_weights_updated(self) -> None:
        """Call after updating weights behind torch's back (fused optimizer): drops cached packs."""
        self._infer_version = None

    def _needs_grad(self) -> bool:
        return torch.is_grad_enabled() and any(p.requires_grad for _, p in self._block_params())

    # -------------------------------------------------------------- reference API
    def forward(self, x: Tensor, return_all_features: bool = False) -> Tensor:
        """CLS path: image -> head(norm(x)[:,0])  (eva_vit_model.py:581-586).  Inference only —
        the CLIPSelf step runs it under no_grad for the frozen teacher (clipself.py:37-38)."""
        if return_all_features:
            raise NotImplementedError("return_all_features is not part of the CLIPSelf hot path")
        if self._needs_grad():
            raise NotImplementedError("gradient through the CLS path is not on the CLIPSelf hot path; "
                                      "call under torch.no_grad()")
        x = self._prep(x)
        return self._infer_engine(x).forward_cls(x)

    def teacher_chunk_schedule(self, R: int):
        """(start, count) pieces in which forward_chunked() walks R crops (one H2D event per piece)."""
        from .tower import chunk_schedule
        return chunk_schedule(R, min(self._infer_engine().chunk_images, max(R, 1)))

    def forward_chunked(self, x: Tensor, events, out: Optional[Tensor] = None) -> Tensor:
        """forward() on a crop tensor that is still being filled by an H2D stream: piece k of
        teacher_chunk_schedule() may be read once events[k] has completed (see training/clipself.py).  events = None:
        the tensor is complete.  `out`: caller-owned [R, embed_dim] f32 result buffer (a stable address keeps the
        native tower's per-(buffers, shape) CUDA graphs hot)."""
        return self._infer_engine().forward_cls(self._prep(x) if events is None else x, out=out, ready_events=events)

    def _prep(self, x: Tensor) -> Tensor:
        if x.dim() == 4:
            input_grid(x, self.cfg)           # shape errors first (ValueError), before any engine is chosen
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        return x.contiguous()

    def encode_dense(self, x: Tensor, keep_shape: bool = True) -> Tensor:
        """[B,3,S,S] -> unit-norm dense map; keep_shape: NCHW view [B,C,h,w] of the NHWC data
        (eva_vit_model.py:588-623), else [B,hw,C]."""
        x = self._prep(x)
        if self._needs_grad():
            dense = _DenseFeatures.apply(self, x, *[p for _, p in self._block_params()])
        else:
            dense = self._infer_engine(x).encode_dense_nograd(x)
        B, h, w, C = dense.shape
        return dense.permute(0, 3, 1, 2) if keep_shape else dense.reshape(B, h * w, C)

    @staticmethod
    def _pack_boxes(normed_boxes: Sequence[Tensor], device) -> tuple:
        counts = [int(b.shape[0]) for b in normed_boxes]
        offsets = torch.tensor([0] + list(np.cumsum(counts)), dtype=torch.int32)
        rois = torch.cat([b.reshape(-1, 4) for b in normed_boxes]).to(device=device, dtype=torch.float32).contiguous()
        return rois, offsets.to(device), int(sum(counts))

    def roi_features_packed(self, x: Tensor, rois: Tensor, img_offsets: Tensor, R: int) -> Tensor:
        """extract_roi_features on already packed boxes (rois [R,4] normalised, image-major)."""
        x = self._prep(x)
        if self._needs_grad():
            return _RoiFeatures.apply(self, x, rois, img_offsets, R, *[p for _, p in self._block_params()])
        dense = self._infer_engine(x).encode_dense_nograd(x)
        return ops.roi_align_fwd(dense, rois, img_offsets, R)[0]

    def extract_roi_features(self, x: Tensor, normed_boxes: Sequence[Tensor], **kwargs) -> Tensor:
        """eva_vit_model.py:625-629 (kwargs such as extract_type are ignored there too)."""
        rois, offsets, R = self._pack_boxes(normed_boxes, x.device)
        return self.roi_features_packed(x, rois, offsets, R)

    def mask_pool(self, x: Tensor, masks: Sequence[Tensor]) -> Tensor:
        """eva_vit_model.py:645-653."""
        counts = [int(m.shape[0]) for m in masks]
        offsets = torch.tensor([0] + list(np.cumsum(counts)), dtype=torch.int32, device=x.device)
        flat_masks = torch.cat([m.float().flatten(-2, -1) for m in masks]).to(x.device).contiguous()
        fmap = self.encode_dense(x, keep_shape=False).contiguous()
        return ops.mask_pool_fwd(fmap, flat_masks, offsets)

    def roi_and_mask_features(self, x: Tensor, normed_boxes: Sequence[Tensor], masks: Sequence[Tensor],
                              normalize: bool = True):
        """RoIAlign features and mask-pooled features of the same boxes from one dense map
        (eva_vit_model.py:625-629 + 645-653 on a single encode_dense)."""
        if self._needs_grad():
            raise NotImplementedError("roi_and_mask_features is an inference path: call it under torch.no_grad()")
        x = self._prep(x)
        rois, offsets, R = self._pack_boxes(normed_boxes, x.device)
        assert [int(m.shape[0]) for m in masks] == [int(b.shape[0]) for b in normed_boxes], "one mask per box"
        dense = self._infer_engine(x).encode_dense_nograd(x)                       # [B,h,w,C]
        B, h, w, C = dense.shape
        roi = ops.roi_align_fwd(dense, rois, offsets, R)[0]
        flat_masks = torch.cat([m.float().flatten(-2, -1) for m in masks]).to(x.device).contiguous()
        pooled = ops.mask_pool_fwd(dense.view(B, h * w, C), flat_masks, offsets)
        if normalize:
            roi, pooled = ops.l2norm_fwd(roi)[0], ops.l2norm_fwd(pooled)[0]
        return roi, pooled


# ----------------------------------------------------------------------------------------------
# CustomCLIP
# ----------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    """A node of the opaque text-tower parameter tree (no forward)."""


def text_param_shapes(text_cfg: dict, embed_dim: int):
    """Names and shapes of the reference's `text.*` state_dict entries (TextTransformer,
    eva_clip/transformer.py:642-742; 149 keys for both EVA02-CLIP configs)."""
    W, L, ctx, vocab = text_cfg["width"], text_cfg["layers"], text_cfg["context_length"], text_cfg["vocab_size"]
    out = [("positional_embedding", (ctx, W)), ("text_projection", (W, embed_dim)), ("token_embedding.weight", (vocab, W))]
    for i in range(L):
        p = f"transformer.resblocks.{i}."
        out += [(p + "ln_1.weight", (W,)), (p + "ln_1.bias", (W,)), (p + "attn.in_proj_weight", (3 * W, W)),
                (p + "attn.in_proj_bias", (3 * W,)), (p + "attn.out_proj.weight", (W, W)), (p + "attn.out_proj.bias", (W,)),
                (p + "ln_2.weight", (W,)), (p + "ln_2.bias", (W,)), (p + "mlp.c_fc.weight", (4 * W, W)),
                (p + "mlp.c_fc.bias", (4 * W,)), (p + "mlp.c_proj.weight", (W, 4 * W)), (p + "mlp.c_proj.bias", (W,))]
    out += [("ln_final.weight", (W,)), ("ln_final.bias", (W,))]
    return out


def _opaque_text_tower(text_cfg: dict, embed_dim: int) -> nn.Module:
    """The frozen text tower as a parameter tree with the reference's key names and shapes.  It is never executed
    on the CLIPSelf path (clipself.py:7-49 touches `visual` only); holding its tensors lets checkpoints round-trip
    all 436 keys (main.py:300-317 saves `model.state_dict()`, eva_clip/factory.py:110-129 / `--resume` load it back,
    strictly in the `--resume` case)."""
    root = _Holder()
    for name, shape in text_param_shapes(text_cfg, embed_dim):
        parts = name.split(".")
        node = root
        for part in parts[:-1]:
            if not hasattr(node, part):
                node.add_module(part, _Holder())
            node = getattr(node, part)
        t = torch.empty(shape)
        leaf = parts[-1]
        if leaf == "bias" or leaf.endswith("_bias"):
            nn.init.zeros_(t)
        elif len(shape) == 1:
            nn.init.ones_(t)
        else:
            nn.init.normal_(t, std=0.02 if "embedding" in name else shape[-1] ** -0.5)
        node.register_parameter(leaf, nn.Parameter(t, requires_grad=False))
    return root


class CustomCLIP(nn.Module):
    """API surface of eva_clip/model.py:272-346.  The text tower is not on the CLIPSelf path
    (frozen and never executed there); it is kept as an opaque parameter holder so checkpoints
    round-trip (`text.*` keys), see DESIGN.md 'out of scope'."""

    def __init__(self, embed_dim: int, vision_cfg: dict, text_cfg: Optional[dict] = None, **unused):
        super().__init__()
        v = dict(vision_cfg)
        width = v["width"]
        self.visual = EVAVisionTransformer(
            img_size=v["image_size"], patch_size=v["patch_size"], num_classes=embed_dim, embed_dim=width,
            depth=v["layers"], num_heads=width // v.get("head_width", 64), mlp_ratio=v.get("mlp_ratio", 4.0),
            pt_hw_seq_len=v.get("pt_hw_seq_len", 16))
        self.text_cfg = dict(text_cfg) if text_cfg else None
        # with a text_cfg: the reference's 149 `text.*` entries as frozen, never-executed parameters
        self.text = _opaque_text_tower(self.text_cfg, embed_dim) if self.text_cfg else None
        self.embed_dim = embed_dim
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))

    def lock_image_tower(self, unlocked_groups=0, freeze_bn_stats=False, **kwargs):
        self.visual.lock(unlocked_groups=unlocked_groups, freeze_bn_stats=freeze_bn_stats)

    def lock_text_tower(self, unlocked_layers: int = 0, freeze_layer_norm: bool = True):
        return None

    def set_grad_checkpointing(self, enable=True):
        self.visual.set_grad_checkpointing(enable)

    def no_weight_decay(self):
        return {"logit_scale"}

    @staticmethod
    def _normalize(x: Tensor, dim: int = -1) -> Tensor:
        if dim in (-1, x.dim() - 1) and x.dim() == 2 and x.dtype == torch.float32 and x.is_cuda and not x.requires_grad:
            return ops.l2norm_fwd(x.contiguous())[0]
        return torch.nn.functional.normalize(x, dim=dim)

    def encode_image(self, image, normalize: bool = False):
        features = self.visual(image)
        return self._normalize(features) if normalize else features

    def encode_text(self, text, normalize: bool = False):
        raise NotImplementedError("the text tower is outside the CLIPSelf hot path (SURVEY.md §2); "
                                  "text embeddings are consumed as precomputed .npy files by the reference's eval")

    def encode_dense(self, image, normalize: bool = False, keep_shape=False):
        features = self.visual.encode_dense(image, keep_shape=keep_shape)
        if normalize:
            features = torch.nn.functional.normalize(features, dim=1 if keep_shape else -1)
        return features

    def encode_pseudo_boxes(self, image, normed_boxes, normalize: bool = False, extract_type="v1"):
        features = self.visual.extract_roi_features(image, normed_boxes, extract_type=extract_type)
        return self._normalize(features) if normalize else features

    def encode_masks(self, image, masks, normalize=True, mask_attn=False):
        pooled = self.visual.mask_pool(image, masks)
        return self._normalize(pooled) if normalize else pooled

    def encode_boxes_and_masks(self, image, normed_boxes, masks, normalize: bool = True):
        """encode_pseudo_boxes + encode_masks on ONE dense pass (the reference's eval loop, zero_shot.py:71-77,
        runs the tower twice for the same images).  Inference only."""
        return self.visual.roi_and_mask_features(image, normed_boxes, masks, normalize)
