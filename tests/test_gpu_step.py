"""Whole CLIPSelf step on the GPU (forward + backward through the plug-in boundary) against the
reference's golden loss / gradients."""
import types

import numpy as np
import pytest
import torch

from oracle import clipself_oracle as O

pytestmark = pytest.mark.gpu

CASES = {"tiny_ragged": (O.CFG_TINY, 3, 5, "proposal", True), "tiny_grid": (O.CFG_TINY, 2, 4, "grid", False),
         "cfg1_b16": (O.CFG_B16, 2, 8, "grid", False), "l14_fwd": (O.CFG_L14_336, 1, 2, "proposal", False),
         "tiny_multires": (O.CFG_TINY, 2, 4, "proposal", True)}
GRAD_TOL = 3e-2        # measured worst 1.2e-2 .. 1.7e-2 (profiles/r02_parity.txt)
DET_SIZE = {"tiny_multires": 160}      # student images at a detector resolution != the tower's own (10x10 grid)


def build_model(ocfg, seed, dev):
    from clipself_b200.model import CustomCLIP
    vis = dict(image_size=ocfg.image_size, layers=ocfg.layers, width=ocfg.width, head_width=64,
               patch_size=ocfg.patch, mlp_ratio=ocfg.hidden / ocfg.width, pt_hw_seq_len=ocfg.pt_seq_len)
    m = CustomCLIP(embed_dim=ocfg.embed_dim, vision_cfg=vis)
    missing, unexpected = m.visual.load_state_dict(O.synth_tower_weights(ocfg, seed), strict=False)
    assert not unexpected and all("rope" in k for k in missing)
    return m.to(dev)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


@pytest.mark.parametrize("tag,host_batch", [("tiny_ragged", True), ("tiny_ragged", False), ("tiny_grid", True),
                                            ("cfg1_b16", True), ("l14_fwd", True), ("tiny_multires", True)])
def test_step_vs_golden(golden, tag, host_batch):
    from clipself_b200.training.clipself import CLIPSelf
    ocfg, B, K, kind, ragged = CASES[tag]
    g = golden(tag)
    seed = int(g["seed"])
    dev = torch.device("cuda")
    student = build_model(ocfg, seed, dev)
    teacher = build_model(ocfg, seed + 1, dev)
    student.lock_image_tower(unlocked_groups=ocfg.layers)
    student.train()
    teacher.eval()
    batch = O.synth_batch(ocfg, B, K, seed + 2, kind=kind, ragged=ragged, det_size=DET_SIZE.get(tag))
    if not host_batch:
        batch = tuple(t.to(dev) for t in batch)
    args = types.SimpleNamespace(multiscale=False, extract_type="v2", cosine_weight=1.0)
    losses, bs, logit_scale = CLIPSelf()(batch, student, teacher, None, dev, None, False, args)
    loss = losses["loss_cosine"]
    assert bs == B
    np.testing.assert_allclose(logit_scale.item(), float(g["logit_scale_exp"]), rtol=1e-6)
    # loss within 1e-3 relative of the reference fp32 value (north_star tolerance)
    print(f"{tag}: loss {loss.item():.6f} ref {float(g['loss']):.6f} (ref bf16-autocast {float(g['ref_autocast_bf16_loss']):.6f})")
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    loss.backward()
    torch.cuda.synchronize()
    names = [str(n) for n in g["grad_names"]]
    worst = 0.0
    checked = 0
    for name, ref_norm in zip(names, g["grad_norms"]):
        if not name.startswith("blocks."):
            continue
        p = dict(student.visual.named_parameters())[name]
        if ref_norm < 0:
            assert p.grad is None, name
            continue
        assert p.grad is not None, name
        gn = p.grad.double().norm().item()
        # bf16 tensor-core backward vs fp32 reference: norms within 6 %, full tensors within 8 % rel-L2
        assert abs(gn - ref_norm) <= 0.06 * ref_norm + 1e-9, (name, gn, ref_norm)
        key = "grad/" + name
        if key in g.files:
            r = rel(p.grad.cpu().numpy(), g[key])
            worst = max(worst, r)
            checked += 1
            assert r <= 0.08, (name, r)
    print(f"{tag}: {checked} gradient tensors compared, worst rel-L2 {worst:.3e}")
    assert checked > 0


@pytest.mark.parametrize("tag", ["tiny_ragged", "tiny_grid", "cfg1_b16", "l14_fwd"])
def test_step_vs_device_arithmetic_oracle(golden, tag):
    """The whole step against the oracle with the kernels' rounding points (oracle/device_arith_oracle.py): loss to 1e-3
    (north_star), dense map to 8e-3 end to end (stage-wise 1e-4: tests/test_gpu_parity_stages.py explains why two bf16
    pipelines cannot meet 1e-3 end to end), every gradient tensor to GRAD_TOL rel-L2 of the oracle's autograd (fp32
    backward at rounded forward values: what is left is the rounding of the backward's tensor-core operands plus the
    forward's rounding flips)."""
    from clipself_b200.training.clipself import CLIPSelf
    from oracle import device_arith_oracle as DA
    ocfg, B, K, kind, ragged = CASES[tag]
    seed = int(golden(tag)["seed"])
    dev = torch.device("cuda")
    student, teacher = build_model(ocfg, seed, dev), build_model(ocfg, seed + 1, dev)
    student.lock_image_tower(unlocked_groups=ocfg.layers)
    student.train()
    teacher.eval()
    batch = O.synth_batch(ocfg, B, K, seed + 2, kind=kind, ragged=ragged)
    args = types.SimpleNamespace(multiscale=False, extract_type="v2", cosine_weight=1.0)
    losses, _, _ = CLIPSelf()(batch, student, teacher, None, dev, None, False, args)
    loss = losses["loss_cosine"]
    loss.backward()
    torch.cuda.synchronize()
    ssd, tsd = O.synth_tower_weights(ocfg, seed), O.synth_tower_weights(ocfg, seed + 1)
    for k, v in ssd.items():
        if k.startswith("blocks."):
            v.requires_grad_(True)
    ref = DA.clipself_step(ssd, tsd, *batch, ocfg)
    ref["loss"].backward()
    dense = student.visual._student._tape.dense.view(ref["dense"].shape).cpu().numpy()
    rd = rel(dense, ref["dense"].detach().numpy())
    rl = abs(loss.item() - ref["loss"].item()) / abs(ref["loss"].item())
    worst, worst_name = 0.0, ""
    for name, p in student.visual.named_parameters():
        if not name.startswith("blocks.") or ssd[name].grad is None:
            continue
        r = rel(p.grad.cpu().numpy(), ssd[name].grad.numpy())
        if r > worst:
            worst, worst_name = r, name
    print(f"{tag}: vs device-arithmetic oracle: dense rel-L2 {rd:.3e}  loss rel {rl:.3e}  worst gradient rel-L2 {worst:.3e} ({worst_name})")
    assert rd <= 8e-3 and rl <= 1e-3
    assert worst <= GRAD_TOL, worst_name


def test_roi_features_and_masks_vs_golden(golden):
    g = golden("tiny_ragged")
    ocfg, B, K, kind, ragged = CASES["tiny_ragged"]
    dev = torch.device("cuda")
    m = build_model(ocfg, int(g["seed"]), dev).eval()
    images, boxes, _ = O.synth_batch(ocfg, B, K, int(g["seed"]) + 2, kind=kind, ragged=ragged)
    rois = [b[b[:, -1] > 0.5, :4].to(dev) for b in boxes]
    with torch.no_grad():
        f = m.encode_pseudo_boxes(images.to(dev), rois, normalize=True)
        counts = [r.shape[0] for r in rois]
        masks = [t.to(dev) for t in torch.split(torch.from_numpy(g["masks"]), counts)]
        mp = m.encode_masks(images.to(dev), masks, normalize=True)
        d = m.encode_dense(images.to(dev), normalize=False, keep_shape=True)
    assert d.shape == (B, ocfg.embed_dim, ocfg.grid, ocfg.grid)
    assert rel(f.cpu().numpy(), g["student_roi_normalized"]) < 1.5e-2
    assert rel(mp.cpu().numpy(), g["mask_pooled"]) < 1.5e-2
    assert rel(d.permute(0, 2, 3, 1).cpu().numpy(), g["dense_nhwc"]) < 1.5e-2


def test_consecutive_ragged_steps_reuse_workspaces(monkeypatch):
    """Several plug-in calls on the same models with different ragged batches (R changes every step, the
    teacher runs in several H2D-streamed chunks): every loss must match the oracle."""
    from clipself_b200.training.clipself import CLIPSelf
    monkeypatch.setenv("CLIPSELF_TEACHER_CHUNK", "4")
    ocfg = O.CFG_TINY
    dev = torch.device("cuda")
    student, teacher = build_model(ocfg, 31, dev), build_model(ocfg, 32, dev)
    student.lock_image_tower(unlocked_groups=ocfg.layers)
    ssd, tsd = O.synth_tower_weights(ocfg, 31), O.synth_tower_weights(ocfg, 32)
    args = types.SimpleNamespace(multiscale=False, extract_type="v2", cosine_weight=1.0)
    method = CLIPSelf()
    for step, (B, K) in enumerate([(3, 7), (2, 3), (3, 7), (4, 6)]):
        batch = O.synth_batch(ocfg, B, K, 500 + step, kind="proposal", ragged=True)
        batch = tuple(t.pin_memory() for t in batch)
        losses, bs, _ = method(batch, student, teacher, None, dev, None, False, args)
        losses["loss_cosine"].backward()
        ref = O.clipself_step(ssd, tsd, *batch, ocfg)["loss"].item()
        got = losses["loss_cosine"].item()
        # workspace-reuse check, not the parity bar (that is test_step_vs_golden): random near-orthogonal tiny
        # features put the loss at ~1.0, where bf16 noise on a handful of boxes is ~1e-3 absolute
        print(f"step {step}: B={B} K={K} loss {got:.6f} oracle {ref:.6f}")
        assert bs == B and abs(got - ref) <= 3e-3 * abs(ref), (step, got, ref)
        student.zero_grad(set_to_none=True)


def test_step_with_an_image_without_boxes():
    """An image whose rows are all padding (valid flag 0) contributes no RoI, no crop and no gradient."""
    from clipself_b200.training.clipself import CLIPSelf
    ocfg = O.CFG_TINY
    dev = torch.device("cuda")
    student, teacher = build_model(ocfg, 41, dev), build_model(ocfg, 42, dev)
    student.lock_image_tower(unlocked_groups=ocfg.layers)
    images, boxes, crops = O.synth_batch(ocfg, 3, 4, 77, kind="grid", ragged=False)
    boxes[1] = 0.0
    args = types.SimpleNamespace(multiscale=False, extract_type="v2", cosine_weight=0.5)
    losses, bs, _ = CLIPSelf()((images, boxes, crops), student, teacher, None, dev, None, False, args)
    losses["loss_cosine"].backward()
    ref = O.clipself_step(O.synth_tower_weights(ocfg, 41), O.synth_tower_weights(ocfg, 42), images, boxes, crops, ocfg,
                          cosine_weight=0.5)["loss"].item()
    assert bs == 3 and abs(losses["loss_cosine"].item() - ref) <= 3e-3 * abs(ref)
    g = student.visual.blocks[0].mlp.w3.weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0


def test_multires_inference_vs_golden(golden):
    """encode_dense / encode_pseudo_boxes / encode_masks / encode_image at a resolution other than the
    tower's own (RoPE regeneration rope.py:179-214, bicubic pos_embed eva_vit_model.py:631-643)."""
    g = golden("tiny_multires")
    ocfg, B, K, kind, ragged = CASES["tiny_multires"]
    dev = torch.device("cuda")
    m = build_model(ocfg, int(g["seed"]), dev).eval()
    images, boxes, _ = O.synth_batch(ocfg, B, K, int(g["seed"]) + 2, kind=kind, ragged=ragged, det_size=160)
    rois = [b[b[:, -1] > 0.5, :4].to(dev) for b in boxes]
    with torch.no_grad():
        d = m.encode_dense(images.to(dev), normalize=False, keep_shape=True)
        f = m.encode_pseudo_boxes(images.to(dev), rois, normalize=True)
        masks = [t.to(dev) for t in torch.split(torch.from_numpy(g["masks"]), [r.shape[0] for r in rois])]
        mp = m.encode_masks(images.to(dev), masks, normalize=True)
        cls = m.encode_image(images.to(dev), normalize=False)
        native = m.encode_dense(images[:, :, :64, :64].contiguous().to(dev), keep_shape=True)   # back to 4x4
    assert d.shape == (B, ocfg.embed_dim, 10, 10) and native.shape == (B, ocfg.embed_dim, 4, 4)
    assert rel(d.permute(0, 2, 3, 1).cpu().numpy(), g["dense_nhwc"]) < 1.5e-2
    assert rel(f.cpu().numpy(), g["student_roi_normalized"]) < 1.5e-2
    assert rel(mp.cpu().numpy(), g["mask_pooled"]) < 1.5e-2
    ref_cls = O.tower_forward_cls(O.synth_tower_weights(ocfg, int(g["seed"])), images, ocfg)
    assert rel(cls.cpu().numpy(), ref_cls.numpy()) < 1.5e-2
    with pytest.raises(ValueError):
        m.encode_dense(torch.zeros(1, 3, 72, 72, device=dev))          # not a multiple of the patch size


def test_b16_student_at_448_vs_oracle():
    """EVA02-B/16 student at 448 px (28x28 grid, 785 tokens: the long-sequence attention kernels) against the
    oracle: loss and a sample of gradients."""
    from clipself_b200.training.clipself import CLIPSelf
    ocfg = O.CFG_B16
    dev = torch.device("cuda")
    student, teacher = build_model(ocfg, 71, dev), build_model(ocfg, 72, dev)
    student.lock_image_tower(unlocked_groups=ocfg.layers)
    student.train()
    teacher.eval()
    batch = O.synth_batch(ocfg, 1, 3, 73, kind="proposal", det_size=448)
    args = types.SimpleNamespace(multiscale=False, extract_type="v2", cosine_weight=1.0)
    losses, _, _ = CLIPSelf()(batch, student, teacher, None, dev, None, False, args)
    losses["loss_cosine"].backward()
    torch.cuda.synchronize()
    ssd, tsd = O.synth_tower_weights(ocfg, 71), O.synth_tower_weights(ocfg, 72)
    watch = ["blocks.0.attn.q_proj.weight", "blocks.5.attn.proj.weight", "blocks.11.mlp.w3.weight", "blocks.3.norm1.weight"]
    for k in watch:
        ssd[k].requires_grad_(True)
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    out = O.clipself_step(ssd, tsd, *batch, ocfg)
    out["loss"].backward()
    print(f"b16@448: loss {losses['loss_cosine'].item():.6f} oracle {out['loss'].item():.6f}")
    assert abs(losses["loss_cosine"].item() - out["loss"].item()) <= 1e-3 * abs(out["loss"].item())
    params = dict(student.visual.named_parameters())
    for k in watch:
        r = rel(params[k].grad.cpu().numpy(), ssd[k].grad.numpy())
        print(f"  {k}: rel-L2 {r:.3e}")
        assert r <= 0.08, (k, r)


@pytest.mark.parametrize("want", [1024, 640])
def test_multiscale_step_vs_oracle(want):
    """args.multiscale (clipself.py:17-27): 1024 px student images, target size drawn with random.choice;
    1024 keeps the full 64x64 grid (4097 tokens), 640 goes through the bilinear resize kernel."""
    import random
    from clipself_b200.training.clipself import CLIPSelf
    ocfg = O.CFG_TINY
    dev = torch.device("cuda")
    student, teacher = build_model(ocfg, 81, dev), build_model(ocfg, 82, dev)
    student.lock_image_tower(unlocked_groups=ocfg.layers)
    student.train()
    teacher.eval()
    seed = next(s for s in range(100) if random.Random(s).choice([320, 640, 896, 1024]) == want)
    batch = O.synth_batch(ocfg, 1, 3, 83, kind="proposal", det_size=1024)
    args = types.SimpleNamespace(multiscale=True, extract_type="v2", cosine_weight=1.0)
    random.seed(seed)
    losses, _, _ = CLIPSelf()(batch, student, teacher, None, dev, None, False, args)
    losses["loss_cosine"].backward()
    torch.cuda.synchronize()
    ssd, tsd = O.synth_tower_weights(ocfg, 81), O.synth_tower_weights(ocfg, 82)
    watch = ["blocks.0.attn.q_proj.weight", "blocks.1.attn.v_proj.weight", "blocks.2.mlp.w3.weight", "blocks.1.norm2.bias"]
    for k in watch:
        ssd[k].requires_grad_(True)
    images = torch.nn.functional.interpolate(batch[0], size=(want, want), mode="bilinear")
    out = O.clipself_step(ssd, tsd, images, batch[1], batch[2], ocfg)
    out["loss"].backward()
    print(f"multiscale {want}: loss {losses['loss_cosine'].item():.6f} oracle {out['loss'].item():.6f}")
    assert abs(losses["loss_cosine"].item() - out["loss"].item()) <= 1e-3 * abs(out["loss"].item())
    params = dict(student.visual.named_parameters())
    for k in watch:
        r = rel(params[k].grad.cpu().numpy(), ssd[k].grad.numpy())
        print(f"  {k}: rel-L2 {r:.3e}")
        assert r <= 0.08, (k, r)


def test_partially_unlocked_tower(golden):
    """--lock-image-unlocked-groups 2 on the 3-block tower: block 0 frozen (no gradient, untouched by the
    optimizer), blocks 1-2 get the same gradients as in the fully unlocked golden run."""
    from clipself_b200.optim import FusedAdamW
    from clipself_b200.training.clipself import CLIPSelf
    g = golden("tiny_ragged")
    ocfg, B, K, kind, ragged = CASES["tiny_ragged"]
    seed = int(g["seed"])
    dev = torch.device("cuda")
    student, teacher = build_model(ocfg, seed, dev), build_model(ocfg, seed + 1, dev)
    student.lock_image_tower(unlocked_groups=2)
    student.train()
    teacher.eval()
    batch = O.synth_batch(ocfg, B, K, seed + 2, kind=kind, ragged=ragged)
    args = types.SimpleNamespace(multiscale=False, extract_type="v2", cosine_weight=1.0)
    losses, _, _ = CLIPSelf()(batch, student, teacher, None, dev, None, False, args)
    assert abs(losses["loss_cosine"].item() - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    losses["loss_cosine"].backward()
    params = dict(student.visual.named_parameters())
    checked = 0
    for name, ref_norm in zip([str(n) for n in g["grad_names"]], g["grad_norms"]):
        if not name.startswith("blocks."):
            continue
        p = params[name]
        if name.startswith("blocks.0.") or ref_norm < 0:
            assert p.grad is None, name
            continue
        assert rel(p.grad.cpu().numpy(), g["grad/" + name]) <= 0.08, name
        checked += 1
    assert checked > 20
    before = {k: v.detach().clone() for k, v in params.items() if k.startswith("blocks.")}
    FusedAdamW(student.visual._student, lr=1e-2, weight_decay=0.1).step()
    torch.cuda.synchronize()
    for k, v in before.items():
        changed = not torch.equal(params[k].detach(), v)
        frozen = k.startswith("blocks.0.") or k in ("blocks.2.attn.q_proj.weight", "blocks.2.attn.k_proj.weight",
                                                   "blocks.2.attn.q_bias")
        assert changed != frozen, k


def test_repeated_backward_does_not_double_gradients():
    """`.grad` of the block parameters are views of the flat gradient buffer the backward overwrites: a second
    step without zero_grad (or with set_to_none=False) must leave the new gradient, not twice it."""
    from clipself_b200.training.clipself import CLIPSelf
    ocfg = O.CFG_TINY
    dev = torch.device("cuda")
    student, teacher = build_model(ocfg, 91, dev), build_model(ocfg, 92, dev)
    student.lock_image_tower(unlocked_groups=ocfg.layers)
    batch = O.synth_batch(ocfg, 2, 4, 93, kind="grid")
    args = types.SimpleNamespace(multiscale=False, extract_type="v2", cosine_weight=1.0)
    method = CLIPSelf()
    norms = []
    for it in range(3):
        if it == 2:
            student.zero_grad(set_to_none=False)
        losses, _, _ = method(batch, student, teacher, None, dev, None, False, args)
        losses["loss_cosine"].backward()
        p = student.visual.blocks[1].mlp.w3.weight
        norms.append((p.grad.double().norm().item(), student.visual.blocks[0].norm1.bias.grad.double().norm().item()))
    for a, b in zip(norms[0], norms[1]):
        assert abs(a - b) <= 1e-3 * a, norms
    for a, b in zip(norms[0], norms[2]):
        assert abs(a - b) <= 1e-3 * a, norms


def test_eval_loop_shared_dense_pass(golden):
    """encode_boxes_and_masks == (encode_pseudo_boxes, encode_masks), and the region-classification loop runs."""
    import types as _t
    from torch.utils.data import DataLoader
    from clipself_b200.data import SyntheticEvalDataset
    from clipself_b200.training import zero_shot
    g = golden("tiny_ragged")
    ocfg, B, K, kind, ragged = CASES["tiny_ragged"]
    dev = torch.device("cuda")
    m = build_model(ocfg, int(g["seed"]), dev).eval()
    images, boxes, _ = O.synth_batch(ocfg, B, K, int(g["seed"]) + 2, kind=kind, ragged=ragged)
    rois = [b[b[:, -1] > 0.5, :4].to(dev) for b in boxes]
    masks = [t.to(dev) for t in torch.split(torch.from_numpy(g["masks"]), [r.shape[0] for r in rois])]
    with torch.no_grad():
        f, mp = m.encode_boxes_and_masks(images.to(dev), rois, masks, normalize=True)
        f2 = m.encode_pseudo_boxes(images.to(dev), rois, normalize=True)
        mp2 = m.encode_masks(images.to(dev), masks, normalize=True)
    torch.testing.assert_close(f, f2, rtol=0, atol=1e-6)
    torch.testing.assert_close(mp, mp2, rtol=0, atol=1e-6)
    assert rel(f.cpu().numpy(), g["student_roi_normalized"]) < 1.5e-2
    ds = SyntheticEvalDataset(ocfg.image_size, ocfg.image_size, 4, num_classes=6, embed_dim=ocfg.embed_dim,
                              downsample_factor=ocfg.patch, length=6, seed=1)
    args = _t.SimpleNamespace(device=dev, distributed=False, image_ave_pool=False, zeroshot_frequency=1, epochs=1)
    res = zero_shot.zero_shot_eval(m, {"val": DataLoader(ds, batch_size=3)}, 1, args)
    assert set(res) == {f"{p}.{t}.macc{k}" for p in ("rois", "crops", "maskpool") for t in ("thing", "stuff") for k in (1, 5)}
    assert all(0.0 <= v <= 1.0 for v in res.values())


def test_training_cli_end_to_end(tmp_path):
    """The reference's launch line (scripts/train_clipself_coco_image_patches_eva_vitb16.sh) on synthetic data, in process:
    --precision amp (GradScaler protocol), --grad-clip-norm, epoch-end ensemble checkpoint with all 436 keys + scaler +
    --save-most-recent, the region-classification eval loop, then --resume of that checkpoint on the image-backed dataset
    type (crops made on the device)."""
    from clipself_b200.training.main import main
    common = ["--batch-size", "4", "--lr", "1e-5", "--wd", "0.1", "--workers", "0", "--model", "EVA02-CLIP-B-16",
              "--pretrained", "eva", "--warmup", "2", "--zeroshot-frequency", "1", "--cache-dir", "", "--log-every-n-steps", "1",
              "--lock-image", "--save-frequency", "1", "--lock-image-unlocked-groups", "12", "--extract-type=v2", "--name", "smoke",
              "--downsample-factor", "16", "--det-image-size", "224", "--alpha", "0.7", "--max-boxes", "6",
              "--train-steps-per-epoch", "3", "--logs", str(tmp_path), "--grad-clip-norm", "5.0"]
    main(common + ["--epochs", "1", "--dataset-type", "synthetic_distill", "--precision", "amp", "--save-most-recent",
                   "--synthetic-eval-classes", "5"])
    ck = tmp_path / "smoke" / "checkpoints"
    ckpt = torch.load(ck / "epoch_1.pt", map_location="cpu")
    assert (ck / "epoch_latest.pt").exists() and ckpt["epoch"] == 1
    assert len(ckpt["state_dict"]) == 436 and "scaler" in ckpt and ckpt["scaler"]["scale"] == 65536.0
    assert int(ckpt["optimizer"]["step"]) == 3
    main(common + ["--epochs", "2", "--dataset-type", "synthetic_images_distill", "--precision", "amp_bf16", "--max-split", "3",
                   "--resume", str(ck / "epoch_1.pt")])
    ckpt2 = torch.load(ck / "epoch_2.pt", map_location="cpu")
    assert ckpt2["epoch"] == 2 and int(ckpt2["optimizer"]["step"]) == 6 and "scaler" not in ckpt2
    k = "visual.blocks.3.mlp.w1.weight"
    assert not torch.equal(ckpt["state_dict"][k], ckpt2["state_dict"][k])           # the resumed epoch trained
    assert torch.equal(ckpt["state_dict"]["text.token_embedding.weight"], ckpt2["state_dict"]["text.token_embedding.weight"])


def test_eval_after_fused_steps_uses_the_updated_weights():
    """the forward-only pack must follow FusedAdamW updates (they bypass torch's version counters)."""
    from clipself_b200.optim import FusedAdamW
    from clipself_b200.training.clipself import CLIPSelf
    ocfg = O.CFG_TINY
    dev = torch.device("cuda")
    student, teacher = build_model(ocfg, 71, dev), build_model(ocfg, 72, dev)
    student.lock_image_tower(unlocked_groups=ocfg.layers)
    batch = O.synth_batch(ocfg, 2, 4, 73, kind="grid")
    args = types.SimpleNamespace(multiscale=False, extract_type="v2", cosine_weight=1.0)
    images = batch[0].to(dev)
    with torch.no_grad():
        before = student.encode_dense(images, keep_shape=False).clone()
    losses, _, _ = CLIPSelf()(batch, student, teacher, None, dev, None, False, args)
    losses["loss_cosine"].backward()
    opt = FusedAdamW(student.visual._student, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    opt.step()
    with torch.no_grad():
        after = student.encode_dense(images, keep_shape=False).clone()
        fresh = build_model(ocfg, 71, dev)
        fresh.load_state_dict(student.state_dict())
        expect = fresh.encode_dense(images, keep_shape=False)
    assert not torch.equal(before, after)
    assert torch.equal(after, expect)
