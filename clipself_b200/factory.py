"""`open_clip.create_model` / `create_model_and_transforms` for the EVA02-CLIP models the reference's
scripts use (src/open_clip/factory.py:111-249 with the 'eva' dispatch :145-149 into
src/open_clip/eva_clip/factory.py:211-355).  Same arguments, same error behaviour for the cases on
the CLIPSelf path; arguments that only concern other backbones are accepted and ignored."""
from __future__ import annotations

import logging
import os
from typing import Optional, Union

import torch

from .configs import get_model_config, list_models  # noqa: F401  (re-export)
from .model import CustomCLIP


def get_cast_dtype(precision: str):
    """eva_clip/model.py:83-89."""
    return {"bf16": torch.bfloat16, "fp16": torch.float16}.get(precision)


def load_state_dict(checkpoint_path: str, map_location="cpu", model_key="model|module|state_dict", is_openai=False,
                    skip_list=()):
    """eva_clip/factory.py:80-106: unwrap, strip 'module.', drop rope buffers (always regenerated)."""
    checkpoint = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
    for mk in model_key.split("|"):
        if isinstance(checkpoint, dict) and mk in checkpoint:
            state_dict = checkpoint[mk]
            break
    else:
        state_dict = checkpoint
    if next(iter(state_dict.items()))[0].startswith("module"):
        state_dict = {k[7:]: v for k, v in state_dict.items()}
    for k in list(state_dict.keys()):
        if k in skip_list or "freqs_cos" in k or "freqs_sin" in k:
            del state_dict[k]
    return state_dict


def resize_evaclip_pos_embed(state_dict, model, interpolation: str = "bicubic", seq_dim=1):
    """eva_clip/utils.py:78-106: when the checkpoint's token grid differs from the model's, bicubically resample the
    patch-position rows of `visual.pos_embed` (CLS row kept) and, as the reference does in the same branch, resample
    `visual.patch_embed.proj.weight` to the model's patch size.  In place on `state_dict`."""
    if "visual.pos_embed" not in state_dict:
        return
    pos = state_dict["visual.pos_embed"]
    emb = pos.shape[-1]
    num_patches = model.visual.patch_embed.num_patches
    num_extra = model.visual.pos_embed.shape[-2] - num_patches
    orig_size = int((pos.shape[-2] - num_extra) ** 0.5)
    new_size = int(num_patches ** 0.5)
    if orig_size == new_size:
        return
    logging.info("Position interpolate from %dx%d to %dx%d" % (orig_size, orig_size, new_size, new_size))
    extra = pos[:, :num_extra]
    tok = pos[:, num_extra:].reshape(-1, orig_size, orig_size, emb).permute(0, 3, 1, 2)
    tok = torch.nn.functional.interpolate(tok.float(), size=(new_size, new_size), mode="bicubic", align_corners=False)
    state_dict["visual.pos_embed"] = torch.cat((extra, tok.permute(0, 2, 3, 1).flatten(1, 2).to(pos.dtype)), dim=1)
    w = state_dict["visual.patch_embed.proj.weight"]
    state_dict["visual.patch_embed.proj.weight"] = torch.nn.functional.interpolate(
        w.float(), size=model.visual.patch_embed.patch_size, mode="bicubic", align_corners=False)


def load_checkpoint(model, checkpoint_path, model_key="model|module|state_dict", strict=False):
    """eva_clip/factory.py:110-129 for EVA checkpoints (`visual.pos_embed` branch)."""
    state_dict = load_state_dict(checkpoint_path, model_key=model_key)
    if "text.logit_scale" in state_dict and hasattr(model, "logit_scale"):
        state_dict["logit_scale"] = state_dict.pop("text.logit_scale")
    resize_evaclip_pos_embed(state_dict, model)
    if model.text is None:           # towers built without a text_cfg hold no text.* entries
        state_dict = {k: v for k, v in state_dict.items() if not k.startswith("text.")}
    incompatible = model.load_state_dict(state_dict, strict=strict)
    logging.info(f"incompatible_keys.missing_keys: {incompatible.missing_keys}")
    return incompatible


def create_model(
        model_name: str,
        pretrained: Optional[str] = None,
        precision: str = "fp32",
        device: Union[str, torch.device] = "cpu",
        jit: bool = False,
        force_quick_gelu: bool = False,
        force_custom_text: bool = False,
        force_patch_dropout: Optional[float] = None,
        force_image_size=None,
        pretrained_image: bool = False,
        pretrained_hf: bool = True,
        cache_dir: Optional[str] = None,
        output_dict: Optional[bool] = None,
        require_pretrained: bool = False,
):
    if jit:
        raise NotImplementedError("jit=True is an OpenAI-checkpoint path, not used by CLIPSelf")
    cfg = get_model_config(model_name)                      # RuntimeError for unknown names, like the reference
    if force_image_size is not None:
        # eva_clip/factory.py:263-265: override the tower's native resolution; a checkpoint's pos_embed is then
        # resampled at load (resize_evaclip_pos_embed)
        size = force_image_size[0] if isinstance(force_image_size, (tuple, list)) else int(force_image_size)
        if size % cfg["vision_cfg"]["patch_size"] != 0:
            raise ValueError(f"force_image_size={size} is not a multiple of the patch size {cfg['vision_cfg']['patch_size']}")
        cfg["vision_cfg"]["image_size"] = size
    if precision in ("bf16", "fp16", "pure_bf16", "pure_fp16"):
        # the reference's pure-bf16 mode dies at torchvision RoIAlign; the runnable
        # bf16 mode is amp_bf16 = f32 master weights + bf16 tensor-core operands, which is what we run.
        raise NotImplementedError(f"precision={precision!r}: use 'amp_bf16' (f32 master weights, bf16 tensor cores)")
    if precision == "fp32" and not getattr(create_model, "_warned_fp32", False):
        # the reference's precision argument only selects the WEIGHT dtype here (eva_clip/factory.py:342-344); the
        # arithmetic is chosen by the caller's autocast context.  This library has one arithmetic: say so once.
        logging.warning("clipself_b200: weights are kept in fp32 (precision='fp32'), but every contraction runs on the "
                        "tensor cores with bf16 operands and fp32 accumulation (the reference's amp_bf16 arithmetic); "
                        "there is no fp32-operand path")
        create_model._warned_fp32 = True
    model = CustomCLIP(embed_dim=cfg["embed_dim"], vision_cfg=cfg["vision_cfg"], text_cfg=cfg["text_cfg"])
    if pretrained and pretrained != "eva":
        raise RuntimeError(f"Pretrained weights ({pretrained}) not found for model {model_name}.")
    if cache_dir:                                           # the scripts pass the checkpoint path here
        if not os.path.exists(cache_dir):
            raise RuntimeError(f"Pretrained weights ({cache_dir}) not found for model {model_name}.")
        load_checkpoint(model, cache_dir)
    elif require_pretrained:
        raise RuntimeError(f"Pretrained weights were required for (model: {model_name}) but not loaded.")
    model.to(device=torch.device(device))
    model.output_dict = bool(output_dict)
    return model


def create_model_and_transforms(model_name: str, pretrained: Optional[str] = None, precision: str = "fp32",
                                device="cpu", jit=False, force_quick_gelu=False, force_custom_text=False,
                                force_patch_dropout=None, force_image_size=None, pretrained_image=False,
                                pretrained_hf=True, image_mean=None, image_std=None, aug_cfg=None,
                                cache_dir: Optional[str] = None, output_dict=None, det_image_size=1024,
                                dataset_type=None):
    """open_clip/factory.py:267-350.  The image transforms belong to the CPU data pipeline, which is
    out of scope; the synthetic dataset does not need them, so None is returned in
    their place."""
    model = create_model(model_name, pretrained, precision, device, jit, force_quick_gelu, force_custom_text,
                         force_patch_dropout, force_image_size, pretrained_image, pretrained_hf, cache_dir, output_dict)
    return model, None, [None, None]
