"""Region-classification evaluation loop of the reference (src/training/zero_shot.py:11-193), host side.

Same contract: `run(model, dataloader, args)` consumes batches
`(images [B,3,S,S], bboxes [B,K,8] = (x0,y0,x1,y1, class, valid, box_size, is_thing), image_crops [B,K,3,s,s],
gt_masks [B,K,h,w], masked_image_crops)` and the dataset's precomputed class embeddings
(`dataloader.dataset.embeddings`, `--embed-path`), and returns the per-box top-5 hit matrices / similarities for
RoIAlign features, crop (CLS) features and mask-pooled features; `zero_shot_eval` turns them into the
thing / stuff mean-accuracy dictionary (`macc_with_is_thing`, zero_shot.py:135-169).

Differences from the reference, all on the device side: the RoI and mask-pooled features of a batch come from
ONE dense pass of the tower (`CustomCLIP.encode_boxes_and_masks`; the reference encodes the images twice,
zero_shot.py:71-77), and the tower kernels are the library's.  The [boxes, C] x [C, classes] logits are a plain
torch matmul: evaluation bookkeeping, not the hot path.
"""
from __future__ import annotations

import logging
from typing import Dict

import torch
import torch.nn.functional as F


def _valid_rows(bboxes_per_image: torch.Tensor) -> torch.Tensor:
    return bboxes_per_image[:, 5] > 0.5                      # zero_shot.py:50


def run(model, dataloader, args):
    cls_embeddings = F.normalize(torch.as_tensor(dataloader.dataset.embeddings).float(), dim=-1).to(args.device)
    module = getattr(model, "module", model)
    correct = {k: [] for k in ("rois", "crops", "maskpool")}
    similarity = {k: [] for k in ("rois", "crops", "maskpool")}
    all_box_sizes, all_is_thing, all_cls_labels = [], [], []
    with torch.no_grad():
        for images, bboxes, image_crops, gt_masks, _masked_image_crops in dataloader:
            images = images.to(args.device, non_blocking=True)
            rois, cls_labels, crops, masks, box_sizes, is_thing = [], [], [], [], [], []
            for boxes_i, crops_i, masks_i in zip(bboxes, image_crops, gt_masks):
                valid = _valid_rows(boxes_i)
                rois.append(boxes_i[valid, :4])
                cls_labels.append(boxes_i[valid, 4])
                crops.append(crops_i[valid])
                masks.append(masks_i[valid])
                box_sizes.append(boxes_i[valid, 6])
                is_thing.append(boxes_i[valid, 7])
            cls_labels = torch.cat(cls_labels).to(torch.long).to(args.device)
            if cls_labels.shape[0] == 0:
                continue
            crops = torch.cat(crops).to(args.device, non_blocking=True)
            all_box_sizes.append(torch.cat(box_sizes).float())
            all_is_thing.append(torch.cat(is_thing))
            roi_features, maskpool_features = module.encode_boxes_and_masks(images, rois, masks, normalize=True)
            if getattr(args, "image_ave_pool", False):
                crop_features = F.normalize(module.visual.encode_dense(crops, keep_shape=True).mean(dim=(-2, -1)), dim=-1)
            else:
                crop_features = module.encode_image(crops, normalize=True)
            for name, feats in (("rois", roi_features), ("crops", crop_features), ("maskpool", maskpool_features)):
                logits = feats.float() @ cls_embeddings.T
                top5 = logits.topk(min(5, logits.shape[1])).indices
                correct[name].append((top5 == cls_labels.view(-1, 1)).cpu())
                similarity[name].append(torch.gather(logits, 1, cls_labels.view(-1, 1))[:, 0].cpu())
            all_cls_labels.append(cls_labels.cpu())
    out = [torch.cat(correct[k]).float() for k in ("rois", "crops", "maskpool")]
    out += [torch.cat(similarity[k]).float() for k in ("rois", "crops", "maskpool")]
    out += [torch.cat(all_box_sizes), torch.cat(all_is_thing), torch.cat(all_cls_labels)]
    if getattr(args, "distributed", False):
        out = [multi_gpu_sync(x) for x in out]
    return tuple(out)


def multi_gpu_sync(x: torch.Tensor) -> torch.Tensor:
    """Concatenate a per-rank tensor over all ranks (ragged lengths allowed), zero_shot.py:128-132."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, x.cpu())
    return torch.cat(parts)


def macc_with_is_thing(correct_matrix: torch.Tensor, is_thing: torch.Tensor, all_cls_labels: torch.Tensor,
                       prefix: str) -> Dict[str, float]:
    """Mean per-class top-1 / top-5 accuracy, separately for thing and stuff classes (zero_shot.py:135-169;
    per-class means are rounded through fp16 exactly like the reference's `.half()`)."""
    def _macc(corrects, cls_labels):
        if cls_labels.numel() == 0:
            return float("nan")
        acc_per_cls = []
        for lb in range(int(cls_labels.min()), int(cls_labels.max()) + 1):
            per_cls = corrects[cls_labels == lb]
            if per_cls.shape[0]:
                acc_per_cls.append(per_cls.mean().half().item())
        return sum(acc_per_cls) / len(acc_per_cls)

    thing, stuff = is_thing > 0, is_thing < 1
    labels = all_cls_labels.long()
    return {f"{prefix}.thing.macc1": _macc(correct_matrix[thing][:, 0], labels[thing]),
            f"{prefix}.thing.macc5": _macc(correct_matrix[thing].sum(-1), labels[thing]),
            f"{prefix}.stuff.macc1": _macc(correct_matrix[stuff][:, 0], labels[stuff]),
            f"{prefix}.stuff.macc5": _macc(correct_matrix[stuff].sum(-1), labels[stuff])}


def zero_shot_eval(model, data, epoch, args) -> Dict[str, float]:
    """zero_shot.py:172-193: runs when a 'val' loader exists and the epoch hits --zeroshot-frequency."""
    if "val" not in data or args.zeroshot_frequency == 0:
        return {}
    if (epoch % args.zeroshot_frequency) != 0 and epoch != args.epochs:
        return {}
    logging.info("Region classifier")
    loader = data["val"].dataloader if hasattr(data["val"], "dataloader") else data["val"]
    c_rois, c_crops, c_mask, _, _, _, _, is_thing, labels = run(model, loader, args)
    results = {}
    results.update(macc_with_is_thing(c_rois, is_thing, labels, "rois"))
    results.update(macc_with_is_thing(c_crops, is_thing, labels, "crops"))
    results.update(macc_with_is_thing(c_mask, is_thing, labels, "maskpool"))
    return results
