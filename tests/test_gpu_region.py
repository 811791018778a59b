"""Region-path kernels (index extraction bit-exact; RoIAlign / mask pool / loss against the oracle)."""
import numpy as np
import pytest
import torch

from oracle import clipself_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from clipself_b200 import _lib
    _lib.require_device()
    return torch.device("cuda")


@pytest.mark.parametrize("B,K,ragged", [(4, 6, True), (64, 32, False), (3, 1, True), (129, 20, True)])
def test_extract_rois_bit_exact(dev, B, K, ragged):
    from clipself_b200 import ops
    _, boxes, _ = O.synth_batch(O.CFG_TINY, B, K, 11, kind="proposal", ragged=ragged, crop_size=16)
    rois_ref, idx_ref = O.extract_rois(boxes)
    rois, crop_index, roi_batch, offsets = ops.extract_rois(boxes.to(dev))
    R = int(offsets[-1])
    assert R == idx_ref.numel()
    assert torch.equal(rois[:R].cpu(), torch.cat(rois_ref))            # bit-exact
    assert crop_index[:R].cpu().tolist() == idx_ref.tolist()
    counts = [r.shape[0] for r in rois_ref]
    assert offsets.cpu().tolist() == [0] + list(np.cumsum(counts))
    assert roi_batch[:R].cpu().tolist() == [b for b, c in enumerate(counts) for _ in range(c)]


def test_extract_rois_all_invalid(dev):
    from clipself_b200 import ops
    boxes = torch.zeros(2, 3, 5, device=dev)
    *_, offsets = ops.extract_rois(boxes)
    assert offsets.cpu().tolist() == [0, 0, 0]


def test_gather_rows(dev):
    from clipself_b200 import ops
    src = torch.randn(10, 3, 8, 8, device=dev)
    idx = torch.tensor([7, 0, 3, 3], device=dev, dtype=torch.int32)
    assert torch.equal(ops.gather_rows(src, idx, 4), src[idx.long()])


@pytest.mark.parametrize("B,H,W,C,K,kind", [(2, 14, 14, 512, 8, "grid"), (3, 4, 4, 64, 5, "proposal"),
                                            (2, 24, 24, 768, 6, "proposal"), (2, 7, 9, 96, 4, "proposal")])
def test_roi_align_fwd_bwd(dev, B, H, W, C, K, kind):
    from clipself_b200 import ops
    tv = pytest.importorskip("torchvision")
    torch.manual_seed(0)
    fmap = torch.randn(B, H, W, C)
    _, boxes, _ = O.synth_batch(O.CFG_TINY, B, K, 5, kind=kind, ragged=(kind == "proposal"), crop_size=16)
    boxes[0, 0, :4] = torch.tensor([0.0, 0.0, 1.0, 1.0])           # whole image
    boxes[-1, 0, :4] = torch.tensor([0.93, 0.9, 1.0, 1.0])         # touches the far edge
    rois, _, _, offsets = ops.extract_rois(boxes.to(dev))
    R = int(offsets[-1])
    out, wy, wx = ops.roi_align_fwd(fmap.to(dev), rois, offsets, R)
    rois_list, _ = O.extract_rois(boxes)
    den = O.denormalize_boxes(rois_list, H, W)
    fm = fmap.clone().requires_grad_(True)
    ref = tv.ops.roi_align(fm.permute(0, 3, 1, 2), den, (1, 1), 1.0, -1, True)[..., 0, 0]
    np.testing.assert_allclose(out.cpu().numpy(), ref.detach().numpy(), rtol=1e-4, atol=2e-6)
    if R <= 12:
        np.testing.assert_allclose(out.cpu().numpy(), O.roi_align_1x1_nhwc(fmap, den).numpy(), rtol=1e-4, atol=2e-6)
    d_out = torch.randn(R, C)
    ref.backward(d_out)
    d_fmap = ops.roi_align_bwd(d_out.to(dev), (B, H, W, C), offsets, R, wy, wx)
    np.testing.assert_allclose(d_fmap.cpu().numpy(), fm.grad.numpy(), rtol=1e-4, atol=2e-6)


def test_mask_pool(dev):
    from clipself_b200 import ops
    torch.manual_seed(1)
    B, h, w, C = 3, 6, 6, 128
    fmap = torch.nn.functional.normalize(torch.randn(B, h, w, C), dim=-1)
    masks = [(torch.rand(n, h, w) > 0.6).float() for n in (2, 0, 3)]
    ref = O.mask_pool(fmap, masks)
    offsets = torch.tensor([0, 2, 2, 5], dtype=torch.int32, device=dev)
    out = ops.mask_pool_fwd(fmap.view(B, h * w, C).to(dev), torch.cat(masks).flatten(1).to(dev), offsets)
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("R,C", [(16, 512), (2048, 512), (37, 64), (1, 768)])
def test_cosine_loss_fwd_bwd(dev, R, C):
    from clipself_b200 import ops
    torch.manual_seed(2)
    s = torch.randn(R, C, requires_grad=True)
    t = torch.randn(R, C) * 3
    ref = O.cosine_loss(s, t, 0.7)
    ref.backward(torch.tensor(1.3))
    loss, stats = ops.cosine_loss_fwd(s.detach().to(dev), t.to(dev), 0.7)
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=2e-6, atol=1e-6)
    d_s = ops.cosine_loss_bwd(s.detach().to(dev), t.to(dev), stats, 0.7, torch.tensor(1.3, device=dev))
    np.testing.assert_allclose(d_s.cpu().numpy(), s.grad.numpy(), rtol=1e-4, atol=1e-8)
    # determinism (fixed-order reduction)
    loss2, _ = ops.cosine_loss_fwd(s.detach().to(dev), t.to(dev), 0.7)
    assert loss2.item() == loss.item()


def test_l2norm_fwd_bwd(dev):
    from clipself_b200 import ops
    torch.manual_seed(3)
    x = torch.randn(777, 512, requires_grad=True)
    y_ref = torch.nn.functional.normalize(x, dim=-1)
    g = torch.randn(777, 512)
    y_ref.backward(g)
    y, inv = ops.l2norm_fwd(x.detach().to(dev))
    np.testing.assert_allclose(y.cpu().numpy(), y_ref.detach().numpy(), rtol=1e-5, atol=1e-7)
    dx = ops.l2norm_bwd(y, inv, g.to(dev))
    np.testing.assert_allclose(dx.cpu().numpy(), x.grad.numpy(), rtol=1e-4, atol=1e-6)


def test_roi_align_degenerate_and_empty_images(dev):
    """Zero-area / inverted boxes pool to 0 (torchvision: empty sampling grid), images without boxes get a
    zero gradient slice, and the oracle agrees."""
    from clipself_b200 import ops
    tv = pytest.importorskip("torchvision")
    torch.manual_seed(4)
    B, H, W, C = 3, 5, 6, 64
    fmap = torch.randn(B, H, W, C)
    boxes = torch.zeros(B, 4, 5)
    boxes[0, 0] = torch.tensor([0.2, 0.2, 0.2, 0.2, 1.0])      # zero area
    boxes[0, 1] = torch.tensor([0.6, 0.7, 0.3, 0.1, 1.0])      # inverted
    boxes[0, 2] = torch.tensor([0.0, 0.0, 1.0, 1.0, 1.0])
    boxes[2, 0] = torch.tensor([0.1, 0.3, 0.9, 0.8, 1.0])      # image 1 has no valid box at all
    rois, _, _, offsets = ops.extract_rois(boxes.to(dev))
    R = int(offsets[-1])
    assert offsets.cpu().tolist() == [0, 3, 3, 4]
    out, wy, wx = ops.roi_align_fwd(fmap.to(dev), rois, offsets, R)
    rois_list, _ = O.extract_rois(boxes)
    den = O.denormalize_boxes(rois_list, H, W)
    ref = tv.ops.roi_align(fmap.permute(0, 3, 1, 2).contiguous(), den, (1, 1), 1.0, -1, True)[..., 0, 0]
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-4, atol=2e-6)
    assert float(out[1].abs().max()) == 0.0
    d_fmap = ops.roi_align_bwd(torch.ones(R, C, device=dev), (B, H, W, C), offsets, R, wy, wx)
    assert float(d_fmap[1].abs().max()) == 0.0 and float(d_fmap[2].abs().max()) > 0


def test_mask_pool_empty_mask_and_large_extract(dev):
    from clipself_b200 import ops
    fmap = torch.randn(2, 9, 32, device=dev)
    masks = torch.zeros(3, 9, device=dev)
    masks[1, 4] = 1.0
    out = ops.mask_pool_fwd(fmap, masks, torch.tensor([0, 2, 3], dtype=torch.int32, device=dev))
    assert float(out[0].abs().max()) == 0.0 and torch.allclose(out[1], fmap[0, 4]) and float(out[2].abs().max()) == 0.0
    _, boxes, _ = O.synth_batch(O.CFG_TINY, 512, 64, 3, kind="proposal", ragged=True, crop_size=2)
    rois_ref, idx_ref = O.extract_rois(boxes)
    rois, crop_index, _, offsets = ops.extract_rois(boxes.to(dev))
    R = int(offsets[-1])
    assert R == idx_ref.numel() and torch.equal(rois[:R].cpu(), torch.cat(rois_ref))
    assert crop_index[:R].cpu().tolist() == idx_ref.tolist()


def test_device_crops_bit_exact_vs_reference_fixture(dev, golden):
    import numpy as np
    from clipself_b200.crops import device_crops, device_det_image
    from oracle import crops_oracle as CO
    g = golden("crops_small")
    img = torch.from_numpy(g["image"]).to(dev)
    got = device_crops(img, g["boxes"], int(g["size"]))
    assert np.array_equal(got.cpu().numpy(), g["crops"])                    # bit-exact f32
    assert np.array_equal(device_det_image(img, int(g["det_size"])).cpu().numpy(), g["det"])
    g2 = golden("crops_coco_like")
    img2 = torch.from_numpy(g2["image"]).to(dev)
    got2 = device_crops(img2, g2["boxes"], int(g2["size"])).cpu().numpy()
    ref2 = np.stack([CO.image_crop(g2["image"], b, int(g2["size"])) for b in g2["boxes"]])
    assert np.array_equal(got2, ref2)


def test_raw_image_batch_step_equals_tensor_batch_step(dev):
    """An image-backed batch (decoded uint8 images + GridDistillDataset boxes, crops made on the device in ONE batched call,
    clipself_b200/crops.py) gives bit-identical student images / crops to the per-image kernels (which the reference
    fixtures pin) and therefore the same loss as the tensor batch built from them."""
    import types
    from clipself_b200.crops import _BatchCropper, device_crops, device_det_image
    from clipself_b200.data import SyntheticImageGridDataset
    from clipself_b200.model import CustomCLIP
    from clipself_b200.training.clipself import CLIPSelf
    from oracle import clipself_oracle as O
    ocfg = O.CFG_TINY
    ds = SyntheticImageGridDataset(det_size=96, crop_size=ocfg.image_size, max_boxes=5, max_split=3, length=8, seed=3, hw=(70, 93))
    samples = [ds[i] for i in range(3)] + [(torch.randint(0, 256, (41, 57, 3), dtype=torch.uint8),) + ds._sample(torch.zeros(41, 57, 3, dtype=torch.uint8), 7)[1:]]
    raw = ds.collate(samples)
    images, crops = _BatchCropper()(raw, dev)
    ref_crops, ref_images = [], []
    for img, px in zip(raw.images_u8, raw.crop_boxes_px):
        ref_images.append(device_det_image(img.to(dev), raw.det_size))
        ref_crops.append(device_crops(img.to(dev), px.tolist(), raw.crop_size))
    assert torch.equal(images, torch.stack(ref_images)) and torch.equal(crops, torch.cat(ref_crops))
    # the same step through the plug-in, once from the raw batch and once from the equivalent tensors
    vis = dict(image_size=ocfg.image_size, layers=ocfg.layers, width=ocfg.width, head_width=64, patch_size=ocfg.patch,
               mlp_ratio=ocfg.hidden / ocfg.width, pt_hw_seq_len=ocfg.pt_seq_len)
    student, teacher = CustomCLIP(ocfg.embed_dim, vis), CustomCLIP(ocfg.embed_dim, vis)
    student.visual.load_state_dict(O.synth_tower_weights(ocfg, 61), strict=False)
    teacher.visual.load_state_dict(O.synth_tower_weights(ocfg, 62), strict=False)
    student, teacher = student.to(dev), teacher.to(dev)
    student.lock_image_tower(unlocked_groups=ocfg.layers)
    args = types.SimpleNamespace(multiscale=False, extract_type="v2", cosine_weight=1.0)
    l_raw = CLIPSelf()(raw, student, teacher, None, dev, None, False, args)[0]["loss_cosine"].item()
    B, K = raw.normed_boxes.shape[:2]
    crops_t = torch.zeros(B, K, 3, raw.crop_size, raw.crop_size, device=dev)
    o = 0
    for b in range(B):
        k = len(raw.crop_boxes_px[b])
        crops_t[b, :k] = crops[o:o + k]
        o += k
    l_tensor = CLIPSelf()((images, raw.normed_boxes.to(dev), crops_t), student, teacher, None, dev, None, False, args)[0]["loss_cosine"].item()
    assert l_raw == l_tensor
