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


def run_grid_sample(tag, h, w, crop_size, det_size, max_anns, choice, seed, crop_scale=1.0):
    """One sample exactly as GridDistillDataset produces it (training/data.py:200-281): the class's own
    _init_boxes / _obtain_image_crops run on a stand-in `self` (the constructor needs COCO annotation files),
    followed by the box re-normalisation lines of __getitem__ (:270-281) copied in spirit below."""
    import random
    import types
    # import-only stand-ins for the annotation libraries (absent here; never called)
    for name in ("pycocotools", "pycocotools.coco", "pycocotools.cocoeval", "pycocotools.mask", "panopticapi", "panopticapi.utils"):
        m = sys.modules.setdefault(name, types.ModuleType(name))
        m.__path__ = []
    sys.modules["pycocotools.coco"].COCO = object
    sys.modules["pycocotools.cocoeval"].COCOeval = object
    for sub in ("coco", "cocoeval", "mask"):
        setattr(sys.modules["pycocotools"], sub, sys.modules["pycocotools." + sub])
    sys.modules["panopticapi"].utils = sys.modules["panopticapi.utils"]
    from open_clip.transform import get_scale
    from training.data import GridDistillDataset
    img = synth_image(h, w, seed)
    pil = Image.fromarray(img)
    me = types.SimpleNamespace(transforms=(det_image_transform(det_size, is_train=False),
                                           image_transform(crop_size, is_train=False, resize_longest_max=True)),
                               max_anns=max_anns, args=types.SimpleNamespace(crop_scale=crop_scale), choices=[choice])
    GridDistillDataset._init_boxes(me)
    random.seed(seed)
    image_crops, boxes = GridDistillDataset._obtain_image_crops(me, pil, choice)
    new_image = me.transforms[0](pil)
    scale = get_scale(pil, new_image)
    boxes_template = torch.zeros(max_anns, 4 + 1)
    crops_template = torch.zeros(max_anns, 3, crop_size, crop_size)
    _, nh, nw = new_image.shape
    boxes[:, :4] *= scale
    boxes[:, [0, 2]] /= nw
    boxes[:, [1, 3]] /= nh
    boxes_template[:boxes.shape[0], :4] = boxes
    boxes_template[:boxes.shape[0], 4] = 1.0
    crops_template[:boxes.shape[0]] = image_crops
    path = os.path.join(HERE, f"{tag}.npz")
    np.savez_compressed(path, image=img, choice=np.array(choice), seed=np.int64(seed), max_anns=np.int64(max_anns),
                        crop_size=np.int64(crop_size), det_size=np.int64(det_size), crop_scale=np.float64(crop_scale),
                        new_image=new_image.numpy(), boxes_template=boxes_template.numpy(), crops_template=crops_template.numpy())
    print(tag, tuple(new_image.shape), tuple(boxes_template.shape), "->", path, f"{os.path.getsize(path) / 1e3:.1f} kB")


if __name__ == "__main__":
    run_grid_sample("grid_sample_small", h=90, w=120, crop_size=32, det_size=64, max_anns=5, choice=(2, 3), seed=700)
    run_grid_sample("grid_sample_scaled", h=75, w=50, crop_size=32, det_size=64, max_anns=8, choice=(3, 2), seed=701,
                    crop_scale=1.5)
    run("crops_small", h=97, w=131, size=32, det_size=48, seed=600, n_boxes=6, store_full=True)
    run("crops_coco_like", h=480, w=640, size=224, det_size=1024, seed=601, n_boxes=8, store_full=False)
