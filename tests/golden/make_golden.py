"""Generate golden fixtures for the CLIPSelf hot path FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, absent on the GPU box):

    python tests/golden/make_golden.py

It imports the unmodified reference (`/root/reference/src`) through the import shims in
`oracle/ref_stubs` (ftfy / timm are not installed), flips every block to the
reference's own math-attention branch (`attn.xattn=False`, eva_vit_model.py:221-246; xformers
is absent), loads the seeded synthetic weights of `oracle.clipself_oracle.synth_tower_weights`
into the reference's `CustomCLIP`, runs the reference's `training.clipself.CLIPSelf.__call__`
and writes what it produced to `tests/golden/*.npz`.  The committed fixtures are what pins the
oracle (tests/test_oracle_vs_golden.py) and the CUDA path (tests/test_gpu_*.py).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = ["/root/reference/src", os.path.join(ROOT, "oracle", "ref_stubs"), ROOT]

import open_clip  # noqa: E402  (the reference)
from open_clip.eva_clip.model import CustomCLIP  # noqa: E402
from training.clipself import CLIPSelf  # noqa: E402

from oracle import clipself_oracle as O  # noqa: E402

TEXT_TINY = dict(context_length=8, vocab_size=64, width=32, heads=2, layers=1)


def build_reference_model(cfg: O.TowerCfg, seed: int, name: str | None):
    """The reference's CustomCLIP with our synthetic visual weights."""
    torch.manual_seed(0)
    if name is not None:
        model = open_clip.create_model(name, "eva", device="cpu", precision="fp32", cache_dir=None)
    else:
        vis = dict(image_size=cfg.image_size, layers=cfg.layers, width=cfg.width,
                   head_width=cfg.head_dim, patch_size=cfg.patch, mlp_ratio=cfg.hidden / cfg.width,
                   eva_model_name="tiny", drop_path_rate=0.0, xattn=False, fusedLN=False, rope=True,
                   pt_hw_seq_len=cfg.pt_seq_len, intp_freq=True, naiveswiglu=True, subln=True)
        os.environ["RoPE"] = "1"
        model = CustomCLIP(embed_dim=cfg.embed_dim, vision_cfg=vis, text_cfg=TEXT_TINY)
    for blk in model.visual.blocks:
        blk.attn.xattn = False
    sd = O.synth_tower_weights(cfg, seed)
    missing, unexpected = model.visual.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("rope" in k for k in missing), missing
    assert model.visual.blocks[0].mlp.w1.weight.shape[0] == cfg.hidden
    return model


def run_case(tag, cfg, name, batch, K, kind, ragged, seed, store_inputs, store_all_grads, tap_blocks=None,
             backward=True, det_size=None):
    student = build_reference_model(cfg, seed, name)
    teacher = build_reference_model(cfg, seed + 1, name)
    student.lock_image_tower(unlocked_groups=cfg.layers)      # main.py:161-166
    student.train()
    teacher.eval()
    images, boxes, crops = O.synth_batch(cfg, batch, K, seed + 2, kind=kind, ragged=ragged, det_size=det_size)
    args = types.SimpleNamespace(multiscale=False, extract_type="v2", cosine_weight=1.0)

    taps = {}
    hooks = []
    for i, blk in enumerate(student.visual.blocks[:-1]):
        if tap_blocks is not None and i not in tap_blocks:
            continue
        hooks.append(blk.register_forward_hook(
            lambda m, inp, out, i=i: taps.__setitem__(f"student_block{i}", out.detach().clone())))
    losses, bs, logit_scale = CLIPSelf()((images, boxes, crops), student, teacher, None, "cpu",
                                         None, False, args)
    for h in hooks:
        h.remove()
    loss = losses["loss_cosine"]
    if backward:
        loss.backward()

    out = dict(seed=np.int64(seed), loss=loss.detach().numpy(),
               logit_scale_exp=logit_scale.detach().numpy(), batch_size=np.int64(bs),
               boxes=boxes.numpy())
    out.update({k: v.numpy() for k, v in taps.items()})
    with torch.no_grad():
        rois = [b[b[:, -1] > 0.5, :4] for b in boxes]
        valid_crops = torch.cat([c[b[:, -1] > 0.5] for b, c in zip(boxes, crops)])
        out["teacher"] = teacher.encode_image(valid_crops, normalize=False).numpy()
        out["student_roi"] = student.encode_pseudo_boxes(images, rois, normalize=False).numpy()
        out["student_roi_normalized"] = student.encode_pseudo_boxes(images, rois, normalize=True).numpy()
        dense = student.encode_dense(images, normalize=False, keep_shape=True)      # NCHW view
        out["dense_nhwc"] = dense.permute(0, 2, 3, 1).contiguous().numpy()
        # masks: the boxes rasterised at feature resolution
        g = (det_size or cfg.image_size) // cfg.patch
        masks = []
        for r in rois:
            m = torch.zeros(r.shape[0], g, g)
            for j, (x0, y0, x1, y1) in enumerate(r.tolist()):
                xa, xb = int(np.floor(x0 * g)), max(int(np.ceil(x1 * g)), int(np.floor(x0 * g)) + 1)
                ya, yb = int(np.floor(y0 * g)), max(int(np.ceil(y1 * g)), int(np.floor(y0 * g)) + 1)
                m[j, ya:yb, xa:xb] = 1.0
            masks.append(m)
        out["masks"] = torch.cat(masks).numpy()
        out["mask_pooled"] = student.encode_masks(images, masks, normalize=True).numpy()
        # yardstick: the reference's own bf16-autocast deviation from its fp32 result
        with torch.autocast("cpu", dtype=torch.bfloat16):
            l16, _, _ = CLIPSelf()((images, boxes, crops), student, teacher, None, "cpu", None, False, args)
            d16 = student.encode_dense(images, normalize=False, keep_shape=True).float()
        out["ref_autocast_bf16_loss"] = l16["loss_cosine"].float().numpy()
        out["ref_autocast_bf16_dense_rel_l2"] = ((d16 - dense).norm() / dense.norm()).numpy()
    if store_inputs:
        out["images"] = images.numpy()
        out["crops"] = crops.numpy()
    else:
        out["images_checksum"] = np.float64(images.double().sum().item())
        out["crops_checksum"] = np.float64(crops.double().sum().item())
    names, norms, sums = [], [], []
    for k, p in (student.visual.named_parameters() if backward else []):
        if "rope" in k:
            continue
        has = p.grad is not None
        names.append(k)
        norms.append(p.grad.double().norm().item() if has else -1.0)     # -1: no grad
        sums.append(p.grad.double().sum().item() if has else 0.0)
        if has and (store_all_grads or p.ndim == 1 and (".11." in k or ".0." in k or ".2." in k)):
            out["grad/" + k] = p.grad.numpy()
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms)
    out["grad_sums"] = np.array(sums)
    path = os.path.join(HERE, f"{tag}.npz")
    np.savez_compressed(path, **out)
    print(tag, "loss", float(loss), "->", path, f"{os.path.getsize(path)/1e6:.2f} MB")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    run_case("tiny_ragged", O.CFG_TINY, None, batch=3, K=5, kind="proposal", ragged=True, seed=100,
             store_inputs=True, store_all_grads=True)
    run_case("tiny_grid", O.CFG_TINY, None, batch=2, K=4, kind="grid", ragged=False, seed=200,
             store_inputs=True, store_all_grads=False)
    # student at a detector resolution != the tower's own (scripts: --det-image-size 1024): 160 px -> 10x10 grid
    # on the 4x4-pretrained tiny tower; exercises the RoPE regeneration and the bicubic pos_embed rescale
    run_case("tiny_multires", O.CFG_TINY, None, batch=2, K=4, kind="proposal", ragged=True, seed=500,
             store_inputs=True, store_all_grads=True, det_size=160)
    # BASELINE.json configs[0]: ViT-B/16, 2x224x224, 8 patch-boxes/img
    run_case("cfg1_b16", O.CFG_B16, "EVA02-CLIP-B-16", batch=2, K=8, kind="grid", ragged=False, seed=300,
             store_inputs=False, store_all_grads=False, tap_blocks=(0, 10))
    # BASELINE.json configs[3]/[4] architecture (EVA02 ViT-L/14 @336), 1 image x 2 boxes, forward + backward
    run_case("l14_fwd", O.CFG_L14_336, "EVA02-CLIP-L-14-336", batch=1, K=2, kind="proposal", ragged=False, seed=400,
             store_inputs=False, store_all_grads=False, tap_blocks=(0,), backward=True)
