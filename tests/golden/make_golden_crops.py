"""Golden fixtures for the crop pipeline, produced by the reference's own transform objects
(`open_clip.transform.image_transform(..., resize_longest_max=True)` = transforms[1] of the distill datasets and
`det_image_transform` = transforms[0], src/open_clip/transform.py:56-191) applied the way
GridDistillDataset._obtain_image_crops does (src/training/data.py:226-245: `transforms[1](image.crop(box))`).
Build container only (needs /root/reference):   python tests/golden/make_golden_crops.py
"""
import os
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = ["/root/reference/src", os.path.join(ROOT, "oracle", "ref_stubs"), ROOT]

from open_clip.transform import det_image_transform, image_transform  # noqa: E402  (the reference)


def synth_image(h, w, seed):
    """Smooth-ish synthetic photo (low-frequency pattern + noise) so the resampler sees real gradients."""
    rng = np.random.Generator(np.random.PCG64(seed))
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    base = np.stack([127 + 100 * np.sin(xx / (7 + c) + yy / (11 + 2 * c)) for c in range(3)], -1)
    return np.clip(base + rng.normal(0, 25, (h, w, 3)), 0, 255).astype(np.uint8)


def boxes_for(h, w, seed, n):
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    xy = rng.random((n, 2)) * np.array([w, h]) * 0.6
    wh = (rng.random((n, 2)) * 0.35 + 0.05) * np.array([w, h])
    b = np.concatenate([xy, xy + wh], 1)
    b[0] = [0.0, 0.0, w, h]                        # the whole image (M = N = 1 grid)
    b[1] = [w / 3.0, 0.0, 2 * w / 3.0, h / 2.0]    # a grid cell with fractional edges
    return b.astype(np.float64)


def run(tag, h, w, size, det_size, seed, n_boxes, store_full):
    img = synth_image(h, w, seed)
    pil = Image.fromarray(img)
    crop_tf = image_transform(size, is_train=False, resize_longest_max=True)
    det_tf = det_image_transform(det_size, is_train=False)
    boxes = boxes_for(h, w, seed, n_boxes)
    crops = torch.stack([crop_tf(pil.crop(tuple(b.tolist()))) for b in boxes]).numpy()
    det = det_tf(pil).numpy()
    out = dict(image=img, boxes=boxes, size=np.int64(size), det_size=np.int64(det_size))
    if store_full:
        out.update(crops=crops, det=det)
    else:                                           # large case: checksums only (float64 sums are order-independent enough: exact compare of bytes hash)
        out.update(crops_sha=np.frombuffer(__import__("hashlib").sha256(crops.tobytes()).digest(), np.uint8),
                   det_sha=np.frombuffer(__import__("hashlib").sha256(det.tobytes()).digest(), np.uint8))
    path = os.path.join(HERE, f"{tag}.npz")
    np.savez_compressed(path, **out)
    print(tag, crops.shape, det.shape, "->", path, f"{os.path.getsize(path) / 1e3:.1f} kB")


if __name__ == "__main__":
    run("crops_small", h=97, w=131, size=32, det_size=48, seed=600, n_boxes=6, store_full=True)
    run("crops_coco_like", h=480, w=640, size=224, det_size=1024, seed=601, n_boxes=8, store_full=False)
