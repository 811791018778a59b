"""Synthetic `(image, boxes[K,5], crops[K,3,s,s])` batches with the contract of the reference's
distillation datasets (src/training/data.py:132,281) — the additive `--dataset-type synthetic_distill`
of SURVEY.md §7 (no COCO on the GPU box).  Box statistics follow SURVEY.md §8d:
  grid      K boxes drawn from the M x M grid templates of GridDistillDataset (data.py:200-224)
  proposal  x0,y0 ~ U(0,.6), w,h ~ U(.1,.4)
"""
from __future__ import annotations

import torch
from torch.utils.data import Dataset


def grid_box_templates(m: int, n: int) -> torch.Tensor:
    """Row-major m x n boxes [x0,y0,x1,y1] from f32 linspace edges (data.py:213-224)."""
    ys = torch.linspace(0, 1, m + 1)
    xs = torch.linspace(0, 1, n + 1)
    x0, y0 = torch.meshgrid(xs[:-1], ys[:-1], indexing="xy")
    x1, y1 = torch.meshgrid(xs[1:], ys[1:], indexing="xy")
    return torch.stack([x0, y0, x1, y1], dim=-1).reshape(m * n, 4)


def synthetic_boxes(K: int, kind: str, g: torch.Generator, ragged: bool = False) -> torch.Tensor:
    boxes = torch.zeros(K, 5)
    if kind == "grid":
        side = 1
        while side * side < K:
            side += 1
        tmpl = grid_box_templates(side, side)
        boxes[:, :4] = tmpl[torch.randperm(tmpl.shape[0], generator=g)[:K]]
    elif kind == "proposal":
        xy = torch.rand(K, 2, generator=g) * 0.6
        wh = torch.rand(K, 2, generator=g) * 0.3 + 0.1
        boxes[:, 0:2] = xy
        boxes[:, 2:4] = xy + wh
    else:
        raise ValueError(f"unknown box kind {kind!r}")
    boxes[:, 4] = 1.0
    if ragged:
        keep = int(torch.randint(1, K + 1, (1,), generator=g))
        boxes[keep:] = 0.0
    return boxes


def synthetic_batch(image_size: int, B: int, K: int, kind: str = "grid", seed: int = 0, ragged: bool = False,
                    crop_size: int | None = None):
    g = torch.Generator().manual_seed(seed)
    s = crop_size or image_size
    images = torch.randn(B, 3, image_size, image_size, generator=g)
    crops = torch.randn(B, K, 3, s, s, generator=g)
    boxes = torch.stack([synthetic_boxes(K, kind, g, ragged) for _ in range(B)])
    return images, boxes, crops


class SyntheticDistillDataset(Dataset):
    """Endless-ish synthetic dataset yielding what GridDistillDataset.__getitem__ returns."""

    def __init__(self, image_size: int, crop_size: int, max_boxes: int, kind: str = "grid", length: int = 4096,
                 seed: int = 0, ragged: bool = True):
        self.image_size, self.crop_size, self.K = image_size, crop_size, max_boxes
        self.kind, self.length, self.seed, self.ragged = kind, length, seed, ragged

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(self.seed * 1000003 + idx)
        image = torch.randn(3, self.image_size, self.image_size, generator=g)
        crops = torch.randn(self.K, 3, self.crop_size, self.crop_size, generator=g)
        boxes = synthetic_boxes(self.K, self.kind, g, self.ragged)
        return image, boxes, crops


class SyntheticEvalDataset(Dataset):
    """Synthetic stand-in for COCOPanopticDataset (training/data.py:284-387): yields
    (image [3,S,S], boxes [K,8] = (x0,y0,x1,y1, class, valid, box_size, is_thing), image_crops [K,3,s,s],
    gt_masks [K, S/df, S/df] = the boxes rasterised at the feature-map resolution, masked_image_crops) and carries
    the class embeddings (`--embed-path`) as `.embeddings` [classes, C]."""

    def __init__(self, image_size: int, crop_size: int, max_boxes: int, num_classes: int, embed_dim: int,
                 downsample_factor: int = 16, length: int = 64, seed: int = 0):
        self.S, self.s, self.K, self.ncls, self.df = image_size, crop_size, max_boxes, num_classes, downsample_factor
        self.length, self.seed = length, seed
        g = torch.Generator().manual_seed(seed)
        self.embeddings = torch.randn(num_classes, embed_dim, generator=g).numpy()

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(self.seed * 1000003 + idx + 1)
        image = torch.randn(3, self.S, self.S, generator=g)
        crops = torch.randn(self.K, 3, self.s, self.s, generator=g)
        b = synthetic_boxes(self.K, "proposal", g, ragged=True)               # [K,5]: box + valid
        boxes = torch.zeros(self.K, 8)
        boxes[:, :4] = b[:, :4]
        boxes[:, 4] = torch.randint(0, self.ncls, (self.K,), generator=g).float()
        boxes[:, 5] = b[:, 4]
        boxes[:, 6] = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) * self.S * self.S
        boxes[:, 7] = (torch.rand(self.K, generator=g) > 0.4).float()
        boxes[b[:, 4] < 0.5] = 0.0
        m = self.S // self.df
        masks = torch.zeros(self.K, m, m)
        for k in range(self.K):
            if boxes[k, 5] > 0.5:
                x0, y0, x1, y1 = (boxes[k, :4] * m).tolist()
                xa, ya = int(x0), int(y0)
                masks[k, ya:max(int(-(-y1 // 1)), ya + 1), xa:max(int(-(-x1 // 1)), xa + 1)] = 1.0
        return image, boxes, crops, masks, crops.clone()


# ------------------------------------------------------------------------------------------------------------------
# Image-backed datasets: they yield DECODED uint8 images + the boxes of GridDistillDataset (training/data.py:135-281);
# the pixel work (K bicubic crops per image, the detector-size student image) happens on the device inside the plug-in
# (clipself_b200/crops.py, csrc/crops.cu), bit-exact with the reference's PIL transforms.
# ------------------------------------------------------------------------------------------------------------------
class _GridSamples(Dataset):
    def __init__(self, det_size: int, crop_size: int, max_boxes: int, max_split: int, crop_scale: float, seed: int):
        from .crops import grid_choices
        self.det_size, self.crop_size, self.max_boxes, self.crop_scale, self.seed = det_size, crop_size, max_boxes, crop_scale, seed
        self.choices = grid_choices(max_split)

    def _sample(self, image_u8: torch.Tensor, idx: int):
        """(image_u8, crop rectangles, boxes_template) for one image: a grid is drawn and its cells shuffled with a
        per-sample seeded RNG (the reference uses the process-global `random`, data.py:230-232, 267-268)."""
        import random
        from .crops import grid_sample_boxes
        rng = random.Random(self.seed * 1000003 + idx)
        m, n = rng.choice(self.choices)
        indices = list(range(m * n))
        rng.shuffle(indices)
        px, template = grid_sample_boxes(int(image_u8.shape[0]), int(image_u8.shape[1]), (m, n), indices, self.max_boxes,
                                         self.det_size, self.crop_scale)
        return image_u8, px, template

    def collate(self, samples):
        from .crops import RawImageBatch
        raw = RawImageBatch([s[0] for s in samples], [s[1] for s in samples], torch.stack([s[2] for s in samples]),
                            self.det_size, self.crop_size)
        raw.prepare()             # pinned blob + descriptor tables: DataLoader-side work, not the step's
        return raw


class SyntheticImageGridDataset(_GridSamples):
    """Random uint8 images of a COCO-like size with GridDistillDataset's boxes (no files needed)."""

    def __init__(self, det_size: int, crop_size: int, max_boxes: int, max_split: int = 6, crop_scale: float = 1.0,
                 length: int = 4096, seed: int = 0, hw=(480, 640)):
        super().__init__(det_size, crop_size, max_boxes, max_split, crop_scale, seed)
        self.length, self.hw = length, hw

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(self.seed * 7919 + idx)
        image = torch.randint(0, 256, (self.hw[0], self.hw[1], 3), generator=g, dtype=torch.uint8)
        return self._sample(image, idx)


class ImageGridDistillDataset(_GridSamples):
    """`--dataset-type grid_distill` (training/data.py:135-281): images named by a COCO-style json (`images[*].file_name`,
    or `coco_url`) under `image_root`, decoded with Pillow to uint8 [H,W,3]; unreadable or tiny images are replaced by
    another index like the reference does (:94-97, 258-261).  `input_filename` may also be a directory: every image file
    in it is used."""

    def __init__(self, input_filename: str, image_root: str, det_size: int, crop_size: int, max_boxes: int, max_split: int = 6,
                 crop_scale: float = 1.0, train_ratio: float = 1.0, seed: int = 0):
        import json
        import os
        import random
        super().__init__(det_size, crop_size, max_boxes, max_split, crop_scale, seed)
        self.image_root = image_root or ""
        if os.path.isdir(input_filename):
            self.image_root = self.image_root or input_filename
            names = sorted(f for f in os.listdir(input_filename) if f.lower().endswith((".jpg", ".jpeg", ".png", ".bmp", ".webp")))
        else:
            with open(input_filename) as f:
                info = json.load(f)["images"]
            names = [im["file_name"] if "file_name" in im else os.path.join(*im["coco_url"].split("/")[-2:]) for im in info]
        if train_ratio < 1.0:
            random.Random(seed).shuffle(names)
            names = names[:int(len(names) * train_ratio)]
        if not names:
            raise RuntimeError(f"no images found in {input_filename}")
        self.names = names

    def __len__(self):
        return len(self.names)

    def read_image(self, name: str):
        import os
        import numpy as np
        from PIL import Image
        try:
            with Image.open(os.path.join(self.image_root, name)) as im:
                arr = np.asarray(im.convert("RGB"))
        except Exception:                                    # noqa: BLE001  (the reference prints and resamples too)
            print(f"Cannot load {os.path.join(self.image_root, name)}", flush=True)
            return None
        if arr.shape[0] < 10 or arr.shape[1] < 10:
            print(f"Invalid image, size {arr.shape[1::-1]}", flush=True)
            return None
        return torch.from_numpy(np.ascontiguousarray(arr))

    def __getitem__(self, idx):
        import random
        image = self.read_image(self.names[idx])
        tries = 0
        while image is None:
            tries += 1
            if tries > 100:
                raise RuntimeError("no readable image found")
            idx = random.choice(range(len(self)))
            image = self.read_image(self.names[idx])
        return self._sample(image, idx)
