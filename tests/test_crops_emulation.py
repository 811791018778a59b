"""The crop kernels' arithmetic, checked on the CPU: tests/host_emul/crops_emul.cpp drives the SAME core functions
the CUDA kernels call (clipself_b200/csrc/crops_core.cuh) in the kernels' index order, built here with g++.
It must reproduce the oracle — and therefore Pillow and the reference's transforms — bit for bit.  (The emulation
is test infrastructure; the product library has no CPU path.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from clipself_b200.crops import OPENAI_DATASET_MEAN, OPENAI_DATASET_STD, crop_descriptors
from oracle import crops_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = tmp_path_factory.mktemp("emul") / "libcrops_emul.so"
    src = os.path.join(ROOT, "tests", "host_emul", "crops_emul.cpp")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", str(out), src])
    lib = C.CDLL(str(out))
    lib.crops_emulate.restype = C.c_int
    lib.crops_emulate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]

    def run(img, boxes, size, center):
        descs, ksize_max, rows_max = crop_descriptors(boxes, size, center)
        K = descs.shape[0]
        out_arr = np.full((K, 3, size, size), np.nan, np.float32)
        img = np.ascontiguousarray(img)
        rc = lib.crops_emulate(img.ctypes.data, img.shape[0], img.shape[1], descs.ctypes.data, K, size, ksize_max, rows_max,
                               (C.c_float * 3)(*OPENAI_DATASET_MEAN), (C.c_float * 3)(*OPENAI_DATASET_STD), out_arr.ctypes.data)
        assert rc == 0, f"emulation reported a violated bound ({rc})"
        return out_arr
    return run


def test_emulation_matches_reference_fixture(emul, golden):
    g = golden("crops_small")
    img, size, det_size = g["image"], int(g["size"]), int(g["det_size"])
    assert np.array_equal(emul(img, g["boxes"], size, True), g["crops"])                      # bit-exact f32
    H, W = img.shape[:2]
    assert np.array_equal(emul(img, np.array([[0.0, 0.0, W, H]]), det_size, False)[0], g["det"])


def test_emulation_matches_oracle_on_random_boxes(emul):
    rng = np.random.default_rng(21)
    img = rng.integers(0, 256, (180, 240, 3), dtype=np.uint8)
    xy = rng.random((24, 2)) * [200, 150] - 10            # some boxes stick out of the image (Image.crop zero fill)
    boxes = np.concatenate([xy, xy + rng.random((24, 2)) * [180, 140] + 1.7], 1)
    boxes[0] = [0, 0, 240, 180]
    boxes[1] = [10.5, 20.5, 11.4, 100.0]                   # 1-pixel wide
    boxes[2] = [30, 30, 30.2, 30.3]                        # rounds to an empty rectangle
    for size in (32, 224):
        got = emul(img, boxes, size, True)
        for b, o in zip(boxes, got):
            rect = O.crop_box_to_rect(b)
            w, h = rect[2] - rect[0], rect[3] - rect[1]
            if w <= 0 or h <= 0 or 0 in O.resized_size(h, w, size):
                # empty rectangle, or a sliver whose resized side rounds to 0 (Pillow raises there): zero canvas
                ref = O.to_tensor_normalize(np.zeros((size, size, 3), np.uint8))
            else:
                ref = O.image_crop(img, b, size)
            assert np.array_equal(o, ref), (b, size)


@pytest.mark.parametrize("tag", ["grid_sample_small", "grid_sample_scaled"])
def test_grid_distill_sample_matches_the_reference_dataset(emul, golden, tag):
    """A whole GridDistillDataset sample (fixture produced by the reference's own _init_boxes / _obtain_image_crops
    and transforms): grid boxes, shuffled selection, crop_scale, crops, detector image and the re-normalised boxes.
    The pixels come from the kernels' host emulation, so this is the device path minus the launches."""
    import random
    import torch
    from clipself_b200.crops import grid_distill_sample
    g = golden(tag)
    M, N = (int(v) for v in g["choice"])
    random.seed(int(g["seed"]))
    indices = list(range(M * N))
    random.shuffle(indices)                                   # what _obtain_image_crops draws (data.py:230-232)
    img = g["image"]

    def crops_fn(image, boxes_px, size):
        return torch.from_numpy(emul(img, np.asarray(boxes_px, np.float64), size, True))

    def det_fn(image, size):
        return torch.from_numpy(emul(img, np.array([[0.0, 0.0, img.shape[1], img.shape[0]]]), size, False)[0])

    new_image, boxes_t, crops_t = grid_distill_sample(torch.from_numpy(img), (M, N), indices, int(g["max_anns"]),
                                                      int(g["det_size"]), int(g["crop_size"]), float(g["crop_scale"]),
                                                      crops_fn=crops_fn, det_fn=det_fn)
    assert np.array_equal(new_image.numpy(), g["new_image"])
    assert np.array_equal(boxes_t.numpy(), g["boxes_template"])               # bit-exact f32 boxes
    assert np.array_equal(crops_t.numpy(), g["crops_template"])
