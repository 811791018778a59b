"""`open_clip.ClipLoss` — API surface named by the north star (src/open_clip/loss.py:66-131).

CLIPSelf itself never calls it (`main.py:271` passes `loss=None` to the method), so it is NOT on the
accelerated path: this is a plain torch module with the reference's constructor arguments, label
caching and local-loss / gathered-loss semantics, kept so code written against open_clip imports and
runs.  horovod is not supported (dead branch in the reference's scripts)."""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F


def gather_features(image_features, text_features, local_loss=False, gather_with_grad=False, rank=0, world_size=1,
                    use_horovod=False):
    if use_horovod:
        raise NotImplementedError("horovod is not supported")
    if gather_with_grad:
        import torch.distributed.nn
        all_image = torch.cat(torch.distributed.nn.all_gather(image_features), dim=0)
        all_text = torch.cat(torch.distributed.nn.all_gather(text_features), dim=0)
        return all_image, all_text
    img = [torch.zeros_like(image_features) for _ in range(world_size)]
    txt = [torch.zeros_like(text_features) for _ in range(world_size)]
    dist.all_gather(img, image_features)
    dist.all_gather(txt, text_features)
    img[rank], txt[rank] = image_features, text_features          # keep the local gradient path
    return torch.cat(img, dim=0), torch.cat(txt, dim=0)


class ClipLoss(nn.Module):
    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1,
                 use_horovod=False):
        super().__init__()
        self.local_loss, self.gather_with_grad, self.cache_labels = local_loss, gather_with_grad, cache_labels
        self.rank, self.world_size, self.use_horovod = rank, world_size, use_horovod
        self.prev_num_logits = 0
        self.labels = {}

    def get_ground_truth(self, device, num_logits) -> torch.Tensor:
        if self.prev_num_logits != num_logits or device not in self.labels:
            labels = torch.arange(num_logits, device=device, dtype=torch.long)
            if self.world_size > 1 and self.local_loss:
                labels = labels + num_logits * self.rank
            if self.cache_labels:
                self.labels[device] = labels
                self.prev_num_logits = num_logits
        else:
            labels = self.labels[device]
        return labels

    def get_logits(self, image_features, text_features, logit_scale):
        if self.world_size > 1:
            all_image, all_text = gather_features(image_features, text_features, self.local_loss, self.gather_with_grad,
                                                  self.rank, self.world_size, self.use_horovod)
            if self.local_loss:
                return logit_scale * image_features @ all_text.T, logit_scale * text_features @ all_image.T
            logits_per_image = logit_scale * all_image @ all_text.T
            return logits_per_image, logits_per_image.T
        return logit_scale * image_features @ text_features.T, logit_scale * text_features @ image_features.T

    def forward(self, image_features, text_features, logit_scale, output_dict=False):
        logits_per_image, logits_per_text = self.get_logits(image_features, text_features, logit_scale)
        labels = self.get_ground_truth(image_features.device, logits_per_image.shape[0])
        total = (F.cross_entropy(logits_per_image, labels) + F.cross_entropy(logits_per_text, labels)) / 2
        return {"contrastive_loss": total} if output_dict else total
