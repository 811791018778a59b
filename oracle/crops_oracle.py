"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy, integer arithmetic) of the reference's crop pipeline:

    GridDistillDataset._obtain_image_crops        src/training/data.py:226-245
        image.crop((x0, y0, x1, y1))              PIL Image.crop: box rounded with Python round(), zero outside
        transforms[1] = image_transform(s, is_train=False, resize_longest_max=True)
                                                  src/open_clip/transform.py:119-133
            ResizeMaxSize(s, BICUBIC, fill=0)     transform.py:26-49  (F.resize on a PIL image = Image.resize)
            ToTensor(), Normalize(mean, std)
    det image: det_image_transform -> ResizeLongest(S) (pad right / bottom)   transform.py:136-191

The arithmetic lives in a third-party dependency that is not under /root/reference: Pillow's
`ImagingResample` (src/libImaging/Resample.c; the reference does not pin Pillow — this container has 12.2.0,
the algorithm below is unchanged since Pillow 3.x): separable two-pass resampling, HORIZONTAL pass first into
a uint8 intermediate, then the VERTICAL pass; per output index the filter support is
`2 * max(in/out, 1)` source pixels around `center = in0 + (i + 0.5) * in/out`, bicubic kernel with a = -0.5,
weights normalised in double precision, converted to 22-bit fixed point (round half away from zero), the
accumulator starts at 2^21 and the result is `clip(acc >> 22, 0, 255)`.

Pinned by tests/test_oracle_crops.py: bit-exact against Pillow itself over random sizes, and against the
reference's own transform objects (imported from /root/reference when present) for whole crops.
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c: 22-bit fixed-point coefficients
OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)       # open_clip/constants.py
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)


def bicubic_filter(x: np.ndarray) -> np.ndarray:
    """Resample.c bicubic_filter, a = -0.5, evaluated in double precision with the C expression order."""
    a = -0.5
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def precompute_coeffs(in_size: int, in0: float, in1: float, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc.
    Returns bounds [out,2] (first source index, tap count) and integer coefficients [out, ksize]."""
    scale = float(np.float32(in1) - np.float32(in0)) / out_size          # (double)(in1 - in0) / outSize, box is float
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)                 # C cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        x = np.arange(xmax, dtype=np.float64)
        w = bicubic_filter((x + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:                                        # sequential double sum, like the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        fx = w * float(1 << PRECISION_BITS)
        ki = np.where(w < 0, np.trunc(-0.5 + fx), np.trunc(0.5 + fx)).astype(np.int64)
        kk[xx, :xmax] = ki
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)      # arithmetic shift, then clamp


def resample_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Image.resize((out_w, out_h), BICUBIC) of a uint8 [H, W, C] image (full-image box), bit-exact."""
    H, W, C = img.shape
    need_h = out_w != W
    need_v = out_h != H
    bounds_h, kk_h, _ = precompute_coeffs(W, 0.0, float(W), out_w)
    bounds_v, kk_v, _ = precompute_coeffs(H, 0.0, float(H), out_h)
    ybox_first = int(bounds_v[0, 0])
    ybox_last = int(bounds_v[out_h - 1, 0] + bounds_v[out_h - 1, 1])
    cur = img
    if need_h:
        bounds_v = bounds_v.copy()
        bounds_v[:, 0] -= ybox_first
        src = img[ybox_first:ybox_last].astype(np.int64)
        tmp = np.empty((ybox_last - ybox_first, out_w, C), np.uint8)
        for xx in range(out_w):
            x0, n = bounds_h[xx]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[:, x0:x0 + n, :], kk_h[xx, :n].astype(np.int64), axes=([1], [0]))
            tmp[:, xx, :] = _clip8(acc)
        cur = tmp
    if need_v:
        src = cur.astype(np.int64)
        out = np.empty((out_h, cur.shape[1], C), np.uint8)
        for yy in range(out_h):
            y0, n = bounds_v[yy]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk_v[yy, :n].astype(np.int64), src[y0:y0 + n], axes=([0], [0]))
            out[yy] = _clip8(acc)
        cur = out
    return cur.copy() if cur is img else cur


def crop_box_to_rect(box: Sequence[float]) -> Tuple[int, int, int, int]:
    """PIL Image.crop: `map(int, map(round, box))` — Python round (half to even)."""
    x0, y0, x1, y1 = (int(round(float(v))) for v in box)
    return x0, y0, x1, y1


def crop_u8(img: np.ndarray, rect: Tuple[int, int, int, int]) -> np.ndarray:
    """PIL crop of a uint8 [H, W, C] image: pixels outside the image are 0."""
    x0, y0, x1, y1 = rect
    H, W, C = img.shape
    out = np.zeros((max(y1 - y0, 0), max(x1 - x0, 0), C), np.uint8)
    sx0, sy0, sx1, sy1 = max(x0, 0), max(y0, 0), min(x1, W), min(y1, H)
    if sx1 > sx0 and sy1 > sy0:
        out[sy0 - y0:sy1 - y0, sx0 - x0:sx1 - x0] = img[sy0:sy1, sx0:sx1]
    return out


def resized_size(h: int, w: int, max_size: int) -> Tuple[int, int]:
    """ResizeMaxSize / ResizeLongest: scale = max_size / float(max(h, w)); Python round() of each side."""
    scale = max_size / float(max(h, w))
    return int(round(h * scale)), int(round(w * scale))


def to_tensor_normalize(img_u8: np.ndarray, mean=OPENAI_DATASET_MEAN, std=OPENAI_DATASET_STD) -> np.ndarray:
    """ToTensor (uint8 -> f32 / 255, HWC -> CHW) then Normalize ((x - mean) / std), all in float32."""
    x = img_u8.astype(np.float32) / np.float32(255.0)
    x = (x - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def resize_pad_u8(img: np.ndarray, max_size: int, center: bool) -> np.ndarray:
    """ResizeMaxSize (center=True: crops, transform.py:26-49) or ResizeLongest (center=False: the detector image,
    transform.py:169-191): bicubic resize to the longest side, zero padding to max_size x max_size."""
    h, w = img.shape[:2]
    nh, nw = resized_size(h, w, max_size)
    res = resample_bicubic_u8(img, nh, nw)
    pad_h, pad_w = max_size - nh, max_size - nw
    top, left = (pad_h // 2, pad_w // 2) if center else (0, 0)
    out = np.zeros((max_size, max_size, img.shape[2]), np.uint8)
    out[top:top + nh, left:left + nw] = res
    return out


def image_crop(img: np.ndarray, box: Sequence[float], size: int, mean=OPENAI_DATASET_MEAN, std=OPENAI_DATASET_STD) -> np.ndarray:
    """One training crop: image.crop(box) -> ResizeMaxSize(size) -> ToTensor -> Normalize  =>  f32 [3, size, size]."""
    return to_tensor_normalize(resize_pad_u8(crop_u8(img, crop_box_to_rect(box)), size, center=True), mean, std)


def det_image(img: np.ndarray, size: int, mean=OPENAI_DATASET_MEAN, std=OPENAI_DATASET_STD) -> np.ndarray:
    """The student's input: ResizeLongest(size) -> ToTensor -> Normalize  =>  f32 [3, size, size]."""
    return to_tensor_normalize(resize_pad_u8(img, size, center=False), mean, std)
