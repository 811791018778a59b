"""CLIPSelf distillation step — drop-in for `training.clipself.CLIPSelf` (src/training/clipself.py:6-49).

Same call signature and return value as the reference plug-in invoked at train.py:91-92:

    losses, batch_size, logit_scale = method(batch, model, dist_model, loss, device, cast_dtype,
                                             distributed, args)

Differences in HOW (not what):
  * index extraction (clipself.py:29-36) is done once on the host when the batch arrives as CPU
    tensors (no per-image device syncs, only valid crops cross PCIe) or by one CUDA kernel when the
    batch is already device resident;
  * teacher forward, student dense forward, RoIAlign, normalise + cosine loss and the whole
    backward run in the sm_100a CUDA library;
  * with `distributed=True` the student gradient is mean-all-reduced once per step over the flat
    gradient buffer (the reference's DDP wrapper is bypassed by `model.module`).
"""
from __future__ import annotations

import random

import torch

from .. import ops
from ..model import cosine_distill_loss


class CLIPSelf:
    def __call__(self, batch, model, dist_model, loss, device, cast_dtype, distributed, args):
        if distributed:
            model = getattr(model, "module", model)
            dist_model = getattr(dist_model, "module", dist_model)
        images, normed_boxes, image_crops = batch       # texts are not paired with images
        device = torch.device(device)
        dtype = cast_dtype if cast_dtype is not None else torch.float32
        B, K = normed_boxes.shape[:2]

        if normed_boxes.device.type == "cpu":
            # host-side, bit-exact: valid = boxes[..., 4] > 0.5, image-major order (clipself.py:29-36)
            boxes32 = normed_boxes.float()
            valid = boxes32[:, :, 4] > 0.5
            counts = valid.sum(1)
            offsets = torch.zeros(B + 1, dtype=torch.int32)
            offsets[1:] = counts.cumsum(0)
            R = int(offsets[-1])
            rois = boxes32[valid][:, :4].contiguous()
            crops = image_crops[valid]                   # only the valid crops cross PCIe
            rois = rois.to(device, non_blocking=True)
            offsets = offsets.to(device, non_blocking=True)
            crops = crops.to(device=device, dtype=dtype, non_blocking=True)
        else:
            rois_all, crop_index, _, offsets = ops.extract_rois(normed_boxes.float().contiguous())
            R = int(offsets[-1])                         # the step's single device->host sync
            rois = rois_all[:R]
            flat = image_crops.reshape(B * K, *image_crops.shape[2:])
            crops = ops.gather_rows(flat.contiguous(), crop_index, R) if R != B * K else flat
            crops = crops.to(dtype)
        images = images.to(device=device, dtype=dtype, non_blocking=True)

        if getattr(args, "multiscale", False):
            cur_h, cur_w = images.shape[2:]
            assert cur_h == cur_w
            if cur_h == 1024:
                tar_sizes = [320, 640, 896, 1024]
            elif cur_h == 896:
                tar_sizes = [336, 448, 672, 896]
            else:
                raise NotImplementedError
            tar_size = random.choice(tar_sizes)
            raise NotImplementedError(f"--multiscale (student at {tar_size}px) needs the variable-resolution "
                                      "tower: SURVEY.md §8f rank 4, not built yet")

        with torch.no_grad():
            teacher_crop_features = dist_model.encode_image(crops, normalize=False)
        model.visual.sync_gradients = bool(distributed)
        student_roi_features = model.visual.roi_features_packed(images, rois, offsets, R)

        loss_cosine = cosine_distill_loss(student_roi_features, teacher_crop_features,
                                          float(getattr(args, "cosine_weight", 1.0)))
        losses = dict(loss_cosine=loss_cosine)
        return losses, len(images), model.logit_scale.exp()
