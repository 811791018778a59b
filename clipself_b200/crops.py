"""Host side of the on-device crop generation (STAGED, see csrc/crops.cu): turns the pixel boxes of
GridDistillDataset._obtain_image_crops (training/data.py:226-245) into the integer crop descriptors the kernels
consume, with exactly the rounding of the reference's CPU path:

  Image.crop(box)           -> x0,y0,x1,y1 = int(round(v))            (Python round: half to even)
  ResizeMaxSize(s) / ResizeLongest(S)  (open_clip/transform.py:26-49, 169-191)
                            -> scale = s / float(max(h, w)); new = round(side * scale); centre / top-left padding
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)      # open_clip/constants.py
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)


def crop_descriptors(boxes_px: np.ndarray, size: int, center: bool = True) -> Tuple[np.ndarray, int, int]:
    """boxes_px [K,4] float (x0,y0,x1,y1 in source pixels) -> (descs int32 [K,8], ksize_max, tmp_rows_max).
    descs rows = (x0, y0, x1, y1, out_w, out_h, pad_left, pad_top); degenerate boxes get out_w = out_h = 0."""
    b = np.rint(np.asarray(boxes_px, np.float64)).astype(np.int64).reshape(-1, 4)       # np.rint: half to even
    w, h = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    ok = (w > 0) & (h > 0)
    longest = np.maximum(np.maximum(w, h), 1).astype(np.float64)
    scale = size / longest
    out_h = np.where(ok, np.rint(h * scale), 0).astype(np.int64)
    out_w = np.where(ok, np.rint(w * scale), 0).astype(np.int64)
    pad_h, pad_w = size - out_h, size - out_w
    top, left = (pad_h // 2, pad_w // 2) if center else (np.zeros_like(pad_h), np.zeros_like(pad_w))
    descs = np.stack([b[:, 0], b[:, 1], b[:, 2], b[:, 3], out_w, out_h, np.where(ok, left, 0), np.where(ok, top, 0)],
                     axis=1).astype(np.int32)
    ratio = 1.0
    if ok.any():
        ratio = max(1.0, float(np.max(w[ok] / np.maximum(out_w[ok], 1))), float(np.max(h[ok] / np.maximum(out_h[ok], 1))))
    ksize_max = int(math.ceil(2.0 * ratio)) * 2 + 1
    tmp_rows_max = int(h[ok].max()) if ok.any() else 0
    return np.ascontiguousarray(descs), ksize_max, tmp_rows_max


def device_crops(image_u8: torch.Tensor, boxes_px: Sequence[Sequence[float]], size: int, center: bool = True,
                 mean=OPENAI_DATASET_MEAN, std=OPENAI_DATASET_STD) -> torch.Tensor:
    """image_u8: uint8 [H,W,3] CUDA tensor (the decoded image); returns f32 [K,3,size,size] crops =
    transforms[1](image.crop(box)) for every box, computed on the device."""
    L.require_device()
    assert image_u8.is_cuda and image_u8.dtype == torch.uint8 and image_u8.dim() == 3 and image_u8.shape[2] == 3
    image_u8 = image_u8.contiguous()
    H, W, _ = image_u8.shape
    descs, ksize_max, tmp_rows_max = crop_descriptors(np.asarray(boxes_px, np.float64), size, center)
    K = descs.shape[0]
    dev = image_u8.device
    out = torch.empty(K, 3, size, size, device=dev, dtype=torch.float32)
    if K == 0:
        return out
    import ctypes as C
    need = C.c_int64(0)
    L.call("cs_crop_workspace_bytes", K, size, ksize_max, tmp_rows_max, C.byref(need))
    ws = torch.empty(int(need.value), device=dev, dtype=torch.uint8)
    d_descs = torch.from_numpy(descs).to(dev)
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    L.call("cs_crop_resize_normalize", image_u8.data_ptr(), H, W, d_descs.data_ptr(), K, size, ksize_max, tmp_rows_max,
           m3, s3, out.data_ptr(), ws.data_ptr(), int(need.value), torch.cuda.current_stream().cuda_stream)
    return out


def device_det_image(image_u8: torch.Tensor, det_size: int, mean=OPENAI_DATASET_MEAN, std=OPENAI_DATASET_STD) -> torch.Tensor:
    """The student's input: det_image_transform = ResizeLongest(det_size) + ToTensor + Normalize -> f32 [3,S,S]."""
    H, W, _ = image_u8.shape
    return device_crops(image_u8, [[0.0, 0.0, float(W), float(H)]], det_size, center=False, mean=mean, std=std)[0]


def grid_choices(max_split: int = 16):
    """GridDistillDataset._init_choices (training/data.py:200-205): the (M, N) grids a sample is drawn from."""
    return [(m, n) for m in range(1, max_split + 1) for n in range((m + 1) // 2, min(m * 2 + 1, max_split + 1))]


def grid_distill_sample(image_u8: torch.Tensor, choice: Tuple[int, int], indices: Sequence[int], max_anns: int,
                        det_size: int, crop_size: int, crop_scale: float = 1.0, crops_fn=None, det_fn=None):
    """One training sample of GridDistillDataset (training/data.py:226-281) from a decoded uint8 [H,W,3] image:
    `indices` is the shuffled order of the M*N grid cells (the dataset draws it with random.shuffle; the caller owns
    the RNG), the first `max_anns` are used.  Returns (image f32 [3,S,S], boxes_template [max_anns,5],
    image_crops_template [max_anns,3,s,s]) — the batch contract of the CLIPSelf plug-in.  The pixels are produced
    by device_crops / device_det_image (crops_fn / det_fn are injection points for the CPU tests)."""
    crops_fn = crops_fn or device_crops
    det_fn = det_fn or device_det_image
    img_h, img_w = int(image_u8.shape[0]), int(image_u8.shape[1])
    px, boxes_template = grid_sample_boxes(img_h, img_w, choice, indices, max_anns, det_size, crop_scale)
    crops = crops_fn(image_u8, px.tolist(), crop_size)
    new_image = det_fn(image_u8, det_size)
    crops_template = torch.zeros(max_anns, 3, crop_size, crop_size, device=crops.device)
    crops_template[:px.shape[0]] = crops
    return new_image, boxes_template, crops_template


# ------------------------------------------------------------------------------------------------------------------
# Whole training batches: the dataset does the (cheap) box arithmetic on the host and ships DECODED uint8 images; the
# plug-in produces the student images and the teacher crops on the device in one batched call each.  Compared with the
# reference's DataLoader (K PIL bicubic resizes per image on the CPU, training/data.py:226-245) the host->device traffic
# of a step drops from B*K float32 crops to B uint8 images.
# ------------------------------------------------------------------------------------------------------------------
class RawImageBatch:
    """One batch of a GridDistillDataset-style dataset before any pixel work.

    images_u8     list of B uint8 [H_i, W_i, 3] CPU tensors (decoded images)
    crop_boxes_px list of B float64 arrays [k_i, 4]: the rectangles `image.crop(...)` would cut (data.py:233-243)
    normed_boxes  f32 [B, max_boxes, 5]: boxes for the student on the padded det canvas + valid flag (data.py:265-277)
    det_size / crop_size: side of the student input / of the teacher crops
    """

    def __init__(self, images_u8, crop_boxes_px, normed_boxes: torch.Tensor, det_size: int, crop_size: int):
        assert len(images_u8) == len(crop_boxes_px) == normed_boxes.shape[0]
        self.images_u8, self.crop_boxes_px, self.normed_boxes = list(images_u8), list(crop_boxes_px), normed_boxes
        self.det_size, self.crop_size = int(det_size), int(crop_size)

        self._prepared = None

    def __len__(self):
        return len(self.images_u8)

    def host_bytes(self) -> int:
        p = self.prepare()
        return sum(int(p[k].numel()) * p[k].element_size() for k in ("blob", "offsets", "hw", "det_image", "det_descs", "crop_image",
                                                                      "crop_descs")) + self.normed_boxes.numel() * 4

    def prepare(self):
        """Host-side staging, done once per batch (by the DataLoader's collate, i.e. off the step's critical path): the
        images packed back to back in ONE pinned uint8 blob and the integer crop descriptors of both kernel calls."""
        if self._prepared is not None:
            return self._prepared
        sizes = [int(i.numel()) for i in self.images_u8]
        blob = torch.empty(max(sum(sizes), 1), dtype=torch.uint8)
        try:
            blob = blob.pin_memory()
        except RuntimeError:                      # no CUDA driver (CPU-only tests): pageable staging
            pass
        offs, o = [], 0
        for img, n in zip(self.images_u8, sizes):
            assert img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 3, "images must be decoded uint8 [H,W,3]"
            blob[o:o + n].copy_(img.reshape(-1))
            offs.append(o)
            o += n
        hw = np.asarray([[int(i.shape[0]), int(i.shape[1])] for i in self.images_u8], np.int32)
        det_descs, dk, dt = [], 1, 0
        crop_descs, crop_img, ck, ct = [np.zeros((0, 8), np.int32)], [np.zeros(0, np.int32)], 1, 0
        for i in range(len(self)):
            H, W = int(hw[i, 0]), int(hw[i, 1])
            d, k1, t1 = crop_descriptors(np.asarray([[0.0, 0.0, float(W), float(H)]]), self.det_size, center=False)
            det_descs.append(d)
            dk, dt = max(dk, k1), max(dt, t1)
            if len(self.crop_boxes_px[i]):
                d, k2, t2 = crop_descriptors(self.crop_boxes_px[i], self.crop_size, center=True)
                crop_descs.append(d)
                crop_img.append(np.full(d.shape[0], i, np.int32))
                ck, ct = max(ck, k2), max(ct, t2)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
        self._prepared = dict(blob=blob[:max(o, 1)], offsets=t(np.asarray(offs, np.int64)), hw=t(hw),
                              det_image=t(np.arange(len(self), dtype=np.int32)), det_descs=t(np.concatenate(det_descs)), det_k=dk, det_t=dt,
                              crop_image=t(np.concatenate(crop_img)), crop_descs=t(np.concatenate(crop_descs)), crop_k=ck, crop_t=ct)
        return self._prepared


def grid_sample_boxes(img_h: int, img_w: int, choice: Tuple[int, int], indices: Sequence[int], max_anns: int, det_size: int,
                      crop_scale: float = 1.0):
    """The box arithmetic of GridDistillDataset.__getitem__ / _obtain_image_crops (data.py:226-281) without the pixels:
    returns (crop rectangles in source pixels [k,4] float64, boxes_template [max_anns,5] f32)."""
    from .data import grid_box_templates
    M, N = choice
    normed = grid_box_templates(M, N)
    idx = list(indices)[:max_anns]
    boxes = normed * torch.tensor([img_w, img_h, img_w, img_h])
    px = []
    for i in idx:
        x0, y0, x1, y1 = boxes[i].tolist()
        if crop_scale > 1.0:
            box_w, box_h = x1 - x0, y1 - y0
            cx, cy = (x1 + x0) / 2, (y1 + y0) / 2
            delta = 0.5 * crop_scale
            x0, y0, x1, y1 = max(cx - box_w * delta, 0), max(cy - box_h * delta, 0), \
                min(cx + box_w * delta, img_w), min(cy + box_h * delta, img_h)
        px.append([x0, y0, x1, y1])
    scale = min(det_size / img_h, det_size / img_w)
    sel = boxes[idx].clone()
    sel[:, :4] *= scale
    sel[:, [0, 2]] /= det_size
    sel[:, [1, 3]] /= det_size
    template = torch.zeros(max_anns, 5)
    template[:len(idx), :4] = sel
    template[:len(idx), 4] = 1.0
    return np.asarray(px, np.float64).reshape(-1, 4), template


class _BatchCropper:
    """Device side of a RawImageBatch: one H2D copy of the packed uint8 images, two batched kernel calls."""

    def __init__(self):
        self._dev = None
        self._bufs = {}      # persistent outputs / workspaces: the tower's CUDA graphs and TMA descriptors are cached per address,
                             # so a fresh allocation every step would re-capture / re-encode them (measured: +36 ms per step)

    def _buffer(self, key, shape, dtype, device):
        t = self._bufs.get(key)
        n = 1
        for d in shape:
            n *= int(d)
        if t is None or t.numel() < n or t.dtype != dtype or t.device != device:
            t = torch.empty(max(n, 1), device=device, dtype=dtype)
            self._bufs[key] = t
        return t[:n].view(*shape)

    def _run(self, blob, tables, desc_image, descs, ksize_max, tmp_rows_max, size, device, tag="crops"):
        import ctypes as C
        K = int(descs.shape[0])
        out = self._buffer((tag, "out"), (K, 3, size, size), torch.float32, device)
        if K == 0:
            return out
        need = C.c_int64(0)
        L.call("cs_crop_workspace_bytes", K, size, ksize_max, tmp_rows_max, C.byref(need))
        ws = self._buffer((tag, "ws"), (int(need.value),), torch.uint8, device)
        d_img, d_desc = desc_image.to(device, non_blocking=True), descs.to(device, non_blocking=True)
        m3, s3 = (C.c_float * 3)(*OPENAI_DATASET_MEAN), (C.c_float * 3)(*OPENAI_DATASET_STD)
        L.call("cs_crop_resize_normalize_batched", blob.data_ptr(), tables[0].data_ptr(), tables[1].data_ptr(),
               d_img.data_ptr(), d_desc.data_ptr(), K, size, ksize_max, tmp_rows_max, m3, s3, out.data_ptr(),
               ws.data_ptr(), int(need.value), torch.cuda.current_stream().cuda_stream)
        return out

    def __call__(self, raw: RawImageBatch, device):
        """-> (student images f32 [B,3,S,S], teacher crops f32 [R,3,s,s] image-major, R = number of crop boxes)."""
        L.require_device()
        p = raw.prepare()
        total = int(p["blob"].numel())
        if self._dev is None or self._dev.numel() < total:
            self._dev = torch.empty(max(total, 1), dtype=torch.uint8, device=device)
        self._dev[:total].copy_(p["blob"], non_blocking=True)
        tables = (p["offsets"].to(device, non_blocking=True), p["hw"].to(device, non_blocking=True))
        images = self._run(self._dev, tables, p["det_image"], p["det_descs"], p["det_k"], p["det_t"], raw.det_size, device, "det")
        crops = self._run(self._dev, tables, p["crop_image"], p["crop_descs"], p["crop_k"], p["crop_t"], raw.crop_size, device, "crops")
        return images, crops

    def cast(self, crops: torch.Tensor, dtype) -> torch.Tensor:
        """crops in the tower's input dtype, in a persistent buffer (same address every step)."""
        if crops.dtype == dtype:
            return crops
        out = self._buffer(("crops", "cast", dtype), tuple(crops.shape), dtype, crops.device)
        out.copy_(crops)
        return out
