"""Pin the crop-pipeline oracle (oracle/crops_oracle.py): bit-exact against Pillow's own resampler and against
fixtures produced by the reference's transform objects (tests/golden/make_golden_crops.py)."""
import hashlib

import numpy as np
import pytest

from oracle import crops_oracle as C


@pytest.mark.parametrize("H,W,oh,ow", [(37, 53, 20, 31), (480, 640, 168, 224), (100, 80, 224, 179), (64, 64, 64, 64),
                                       (33, 200, 37, 224), (511, 77, 224, 34), (5, 7, 3, 4), (17, 19, 224, 224)])
def test_resample_is_bit_exact_against_pillow(H, W, oh, ow):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(H * 1000 + W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    assert np.array_equal(C.resample_bicubic_u8(img, oh, ow), ref)


def test_crop_rounding_and_zero_fill_like_pillow():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (40, 50, 3), dtype=np.uint8)
    for box in [(0.5, 1.5, 20.5, 30.5), (2.5, 3.5, 10.49, 12.51), (-3.2, -1.0, 20.0, 45.7), (10.0, 10.0, 60.0, 20.0)]:
        ref = np.asarray(Image.fromarray(img).crop(box))
        got = C.crop_u8(img, C.crop_box_to_rect(box))
        assert got.shape == ref.shape and np.array_equal(got, ref), box


def test_small_fixture_from_the_reference_transforms(golden):
    g = golden("crops_small")
    img, size, det_size = g["image"], int(g["size"]), int(g["det_size"])
    for box, ref in zip(g["boxes"], g["crops"]):
        got = C.image_crop(img, box, size)
        assert got.dtype == np.float32 and np.array_equal(got, ref)            # bit-exact float32
    assert np.array_equal(C.det_image(img, det_size), g["det"])


def test_coco_sized_fixture_checksums(golden):
    """480x640 image, 224 px crops, 1024 px detector image (the published recipe's sizes): SHA-256 of the bytes."""
    g = golden("crops_coco_like")
    img, size, det_size = g["image"], int(g["size"]), int(g["det_size"])
    crops = np.stack([C.image_crop(img, b, size) for b in g["boxes"]])
    assert hashlib.sha256(crops.tobytes()).digest() == g["crops_sha"].tobytes()
    assert hashlib.sha256(C.det_image(img, det_size).tobytes()).digest() == g["det_sha"].tobytes()
