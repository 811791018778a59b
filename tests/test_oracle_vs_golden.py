"""Pin the CPU oracle against outputs of the reference itself (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference).  fp32 on both sides; tolerances cover
summation-order differences only."""
import numpy as np
import pytest
import torch

from oracle import clipself_oracle as O

CASES = {
    "tiny_ragged": (O.CFG_TINY, 3, 5, "proposal", True),
    "tiny_grid": (O.CFG_TINY, 2, 4, "grid", False),
    "cfg1_b16": (O.CFG_B16, 2, 8, "grid", False),
    "l14_fwd": (O.CFG_L14_336, 1, 2, "proposal", False),
    "tiny_multires": (O.CFG_TINY, 2, 4, "proposal", True),      # student images at 160 px (10x10 grid)
}
DET_SIZE = {"tiny_multires": 160}


def _run(golden, tag, need_grad):
    cfg, B, K, kind, ragged = CASES[tag]
    g = golden(tag)
    seed = int(g["seed"])
    images, boxes, crops = O.synth_batch(cfg, B, K, seed + 2, kind=kind, ragged=ragged, det_size=DET_SIZE.get(tag))
    assert np.array_equal(boxes.numpy(), g["boxes"])                    # bit-exact boxes
    if "images" in g.files:
        assert np.array_equal(images.numpy(), g["images"])
        assert np.array_equal(crops.numpy(), g["crops"])
    else:
        assert images.double().sum().item() == float(g["images_checksum"])
        assert crops.double().sum().item() == float(g["crops_checksum"])
    ssd = O.synth_tower_weights(cfg, seed)
    tsd = O.synth_tower_weights(cfg, seed + 1)
    if need_grad:
        for k, v in ssd.items():
            if k.startswith("blocks."):
                v.requires_grad_(True)
    out = O.clipself_step(ssd, tsd, images, boxes, crops, cfg)
    return g, out, ssd


@pytest.mark.parametrize("tag", ["tiny_ragged", "tiny_grid", "tiny_multires"])
def test_forward_tiny(golden, tag):
    g, out, _ = _run(golden, tag, need_grad=False)
    np.testing.assert_allclose(out["teacher"].numpy(), g["teacher"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(out["dense"].numpy(), g["dense_nhwc"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(out["student_roi"].numpy(), g["student_roi"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(out["loss"].item(), float(g["loss"]), rtol=1e-5)


@pytest.mark.parametrize("tag", ["tiny_ragged", "tiny_multires"])
def test_backward_tiny(golden, tag):
    g, out, ssd = _run(golden, tag, need_grad=True)
    out["loss"].backward()
    names = [str(n) for n in g["grad_names"]]
    for name, norm in zip(names, g["grad_norms"]):
        t = ssd[name]
        if not name.startswith("blocks."):
            continue
        if norm < 0:                                   # reference: grad is None
            assert t.grad is None or float(t.grad.abs().sum()) == 0.0, name
            continue
        ref = g["grad/" + name]
        np.testing.assert_allclose(t.grad.numpy(), ref, rtol=2e-3, atol=1e-7 + 2e-4 * np.abs(ref).max(),
                                   err_msg=name)


def test_gradless_params(golden):
    """Last block q/k projections never receive gradient (SURVEY §7 'hard parts', a15)."""
    g = golden("tiny_ragged")
    last = O.CFG_TINY.layers - 1
    none = {str(n) for n, v in zip(g["grad_names"], g["grad_norms"]) if v < 0 and str(n).startswith("blocks.")}
    assert none == {f"blocks.{last}.attn.q_proj.weight", f"blocks.{last}.attn.k_proj.weight",
                    f"blocks.{last}.attn.q_bias"}


def test_mask_pool(golden):
    g, out, _ = _run(golden, "tiny_ragged", need_grad=False)
    B = out["dense"].shape[0]
    boxes = torch.from_numpy(g["boxes"])
    counts = [int((boxes[b, :, 4] > 0.5).sum()) for b in range(B)]
    masks = torch.split(torch.from_numpy(g["masks"]), counts)
    pooled = torch.nn.functional.normalize(O.mask_pool(out["dense"], masks), dim=-1)
    np.testing.assert_allclose(pooled.numpy(), g["mask_pooled"], rtol=1e-4, atol=2e-6)


def test_forward_cfg1_b16(golden):
    """BASELINE.json configs[0] (ViT-B/16, 2x224^2, 8 patch boxes / image, fp32 CPU)."""
    g, out, _ = _run(golden, "cfg1_b16", need_grad=False)
    np.testing.assert_allclose(out["teacher"].numpy(), g["teacher"], rtol=2e-4, atol=5e-5)
    np.testing.assert_allclose(out["dense"].numpy(), g["dense_nhwc"], rtol=2e-4, atol=5e-6)
    np.testing.assert_allclose(out["student_roi"].numpy(), g["student_roi"], rtol=2e-4, atol=5e-6)
    np.testing.assert_allclose(out["loss"].item(), float(g["loss"]), rtol=1e-5)


def test_extract_rois_bit_exact():
    _, boxes, _ = O.synth_batch(O.CFG_TINY, 4, 6, 7, kind="proposal", ragged=True)
    rois, idx = O.extract_rois(boxes)
    ref_idx = []
    for b in range(4):
        valid = boxes[b, :, -1] > 0.5
        assert torch.equal(rois[b], boxes[b][valid][:, :4])
        ref_idx += [b * 6 + int(k) for k in torch.nonzero(valid).flatten()]
    assert idx.tolist() == ref_idx


def test_roi_align_matches_torchvision():
    """The restated RoIAlign against the dependency the reference calls (eva_vit_model.py:628)."""
    tv = pytest.importorskip("torchvision")
    torch.manual_seed(0)
    f = torch.randn(2, 7, 9, 16)
    boxes = [torch.tensor([[0.0, 0.0, 9.0, 7.0], [1.3, 2.2, 1.9, 2.5], [4.0, 3.0, 8.999, 6.5],
                           [8.5, 6.5, 9.0, 7.0]]),
             torch.tensor([[2.0, 1.0, 2.0, 1.0], [0.2, 0.1, 6.7, 5.3]])]
    ours = O.roi_align_1x1_nhwc(f, boxes)
    ref = tv.ops.roi_align(f.permute(0, 3, 1, 2).contiguous(), boxes, (1, 1), 1.0, -1, True)[..., 0, 0]
    np.testing.assert_allclose(ours.numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)


def test_forward_l14_336(golden):
    """EVA02 ViT-L/14 @336 (BASELINE.json configs[3]/[4] architecture): hidden 2730, patch 14, 577 tokens."""
    g, out, _ = _run(golden, "l14_fwd", need_grad=False)
    np.testing.assert_allclose(out["teacher"].numpy(), g["teacher"], rtol=5e-4, atol=1e-4)
    np.testing.assert_allclose(out["dense"].numpy(), g["dense_nhwc"], rtol=5e-4, atol=1e-5)
    np.testing.assert_allclose(out["loss"].item(), float(g["loss"]), rtol=2e-5)


# ----------------------------------------------------------------------------------------------------------------
# device-arithmetic oracle (oracle/device_arith_oracle.py): anchored on the fp32 oracle above
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fold", [True, False])
def test_device_arith_oracle_reduces_to_the_fp32_oracle(fold):
    """With every rounding point disabled the device-arithmetic restatement (folded and explicit LayerNorm forms) is the
    fp32 oracle, which the reference fixtures pin."""
    from oracle import device_arith_oracle as DA
    cfg = O.CFG_TINY
    sd = O.synth_tower_weights(cfg, 5)
    images, boxes, crops = O.synth_batch(cfg, 2, 4, 6, kind="proposal", ragged=True)
    with torch.no_grad():
        a = DA.tower_forward_cls(sd, crops.flatten(0, 1), cfg, fold=fold, exact=True)
        b = O.tower_forward_cls(sd, crops.flatten(0, 1), cfg)
        torch.testing.assert_close(a, b, rtol=2e-4, atol=2e-6)
        a = DA.tower_encode_dense(sd, images, cfg, fold=fold, exact=True)
        b = O.tower_encode_dense(sd, images, cfg)
        torch.testing.assert_close(a, b, rtol=2e-4, atol=2e-6)


def test_device_arith_oracle_rounding_is_the_bf16_noise():
    """Rounding on: the deviation from fp32 is of the size of the reference's own bf16-autocast deviation (the yardstick
    stored in the fixtures), i.e. the rounding points are neither missing nor doubled."""
    from oracle import device_arith_oracle as DA
    cfg = O.CFG_TINY
    sd = O.synth_tower_weights(cfg, 5)
    images, _, _ = O.synth_batch(cfg, 2, 4, 6, kind="proposal", ragged=True)
    with torch.no_grad():
        ref = O.tower_encode_dense(sd, images, cfg)
        for fold in (True, False):
            got = DA.tower_encode_dense(sd, images, cfg, fold=fold)
            r = ((got - ref).norm() / ref.norm()).item()
            assert 5e-4 < r < 2e-2, (fold, r)


def test_device_arith_oracle_step_gradients_flow():
    from oracle import device_arith_oracle as DA
    cfg = O.CFG_TINY
    ssd, tsd = O.synth_tower_weights(cfg, 7), O.synth_tower_weights(cfg, 8)
    for k, v in ssd.items():
        if k.startswith("blocks."):
            v.requires_grad_(True)
    batch = O.synth_batch(cfg, 2, 4, 9, kind="grid")
    out = DA.clipself_step(ssd, tsd, *batch, cfg)
    out["loss"].backward()
    ref = O.clipself_step({k: v.detach() for k, v in ssd.items()}, tsd, *batch, cfg)
    assert abs(out["loss"].item() - ref["loss"].item()) < 5e-3 * abs(ref["loss"].item())
    assert ssd["blocks.0.mlp.w3.weight"].grad.abs().sum() > 0
    assert ssd[f"blocks.{cfg.layers - 1}.attn.q_proj.weight"].grad is None      # grad-less in the reference too


@pytest.mark.parametrize("tag", ["tiny_grid", "cfg1_b16"])
def test_cls_only_last_block_is_the_reference_teacher(golden, tag):
    """The device runs the teacher's last block on the CLS row only (eva_vit_model.py:565-569 reads nothing else).  The
    restatement of that form must reproduce the reference's teacher features of the fixture and the full-form oracle."""
    from oracle import clipself_oracle as O
    cases = {"tiny_grid": (O.CFG_TINY, 2, 4, "grid"), "cfg1_b16": (O.CFG_B16, 2, 8, "grid")}
    ocfg, B, K, kind = cases[tag]
    g = golden(tag)
    seed = int(g["seed"])
    _, boxes, crops = O.synth_batch(ocfg, B, K, seed + 2, kind=kind, ragged=False)
    _, idx = O.extract_rois(boxes)
    tc = crops.flatten(0, 1)[idx]
    sd = O.synth_tower_weights(ocfg, seed + 1)
    with torch.no_grad():
        tail = O.tower_forward_cls_tail(sd, tc, ocfg)
        full = O.tower_forward_cls(sd, tc, ocfg)
    torch.testing.assert_close(tail, full, rtol=1e-4, atol=1e-5)
    ref = torch.from_numpy(np.asarray(g["teacher"]))
    assert ((tail - ref).norm() / ref.norm()).item() < 2e-3
