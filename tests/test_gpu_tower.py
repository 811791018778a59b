"""Tower forward on the GPU against the reference's golden outputs and the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import clipself_oracle as O

pytestmark = pytest.mark.gpu


def _cfg(o):
    from clipself_b200.tower import TowerCfg
    return TowerCfg(image_size=o.image_size, patch=o.patch, width=o.width, heads=o.heads, layers=o.layers,
                    hidden=o.hidden, embed_dim=o.embed_dim, pt_seq_len=o.pt_seq_len, ln_eps=o.ln_eps)


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


CASES = {"tiny_ragged": (O.CFG_TINY, 3, 5, "proposal", True), "tiny_grid": (O.CFG_TINY, 2, 4, "grid", False),
         "cfg1_b16": (O.CFG_B16, 2, 8, "grid", False), "l14_fwd": (O.CFG_L14_336, 1, 2, "proposal", False)}


@pytest.mark.parametrize("tag", ["tiny_ragged", "tiny_grid", "cfg1_b16", "l14_fwd"])
def test_tower_forward_vs_golden(golden, tag):
    """bf16 tensor-core path vs the reference's fp32 outputs.  Tolerance: rel-L2 <= 1.5e-2 on
    feature maps (the reference's own bf16-autocast deviation is stored in the fixture as the
    yardstick: ours must not be worse than 1.5x theirs + 2e-3)."""
    from clipself_b200.tower import TowerEngine
    ocfg, B, K, kind, ragged = CASES[tag]
    g = golden(tag)
    seed = int(g["seed"])
    dev = torch.device("cuda")
    images, boxes, crops = O.synth_batch(ocfg, B, K, seed + 2, kind=kind, ragged=ragged)
    cfg = _cfg(ocfg)
    student = TowerEngine(cfg, O.synth_tower_weights(ocfg, seed), dev)
    teacher = TowerEngine(cfg, O.synth_tower_weights(ocfg, seed + 1), dev)
    _, idx = O.extract_rois(boxes)
    tc = crops.flatten(0, 1)[idx].to(dev)
    t = teacher.forward_cls(tc).cpu().numpy()
    d = student.encode_dense_nograd(images.to(dev)).cpu().numpy()
    yard = float(g["ref_autocast_bf16_dense_rel_l2"])
    rt, rd = _rel(t, g["teacher"]), _rel(d, g["dense_nhwc"])
    print(f"{tag}: teacher rel-L2 {rt:.3e}  dense rel-L2 {rd:.3e}  (reference bf16-autocast dense yardstick {yard:.3e})")
    assert rd <= 1.5 * yard + 2e-3
    assert rt <= 2.5e-2


@pytest.mark.parametrize("tag", ["tiny_ragged", "tiny_grid", "cfg1_b16", "l14_fwd"])
def test_tower_forward_vs_device_arithmetic_oracle(golden, tag):
    """End to end against the oracle that rounds to bf16 exactly where the kernels do (oracle/device_arith_oracle.py).
    Stage by stage the kernels match it to 1e-4 (tests/test_gpu_parity_stages.py holds the 1e-3 of north_star there);
    end to end two non-bit-identical bf16 pipelines settle at a few 1e-3 (rounding flips, see that file's header), so the
    bound here is 8e-3 — still below the reference's own bf16-autocast deviation (1.1e-2) that the fp32 test above uses."""
    from clipself_b200.tower import TowerEngine
    from oracle import device_arith_oracle as DA
    ocfg, B, K, kind, ragged = CASES[tag]
    seed = int(golden(tag)["seed"])
    dev = torch.device("cuda")
    images, boxes, crops = O.synth_batch(ocfg, B, K, seed + 2, kind=kind, ragged=ragged)
    cfg = _cfg(ocfg)
    ssd, tsd = O.synth_tower_weights(ocfg, seed), O.synth_tower_weights(ocfg, seed + 1)
    student, teacher = TowerEngine(cfg, ssd, dev), TowerEngine(cfg, tsd, dev)
    assert teacher.fold_norm, "the folded pipeline is the default for frozen towers"
    _, idx = O.extract_rois(boxes)
    tc = crops.flatten(0, 1)[idx]
    t = teacher.forward_cls(tc.to(dev)).cpu().numpy()
    d = student.encode_dense_nograd(images.to(dev)).cpu().numpy()
    with torch.no_grad():
        t_ref = DA.tower_forward_cls(tsd, tc, ocfg, fold=True).numpy()
        d_ref = DA.tower_encode_dense(ssd, images, ocfg, fold=True).numpy()
    rt, rd = _rel(t, t_ref), _rel(d, d_ref)
    print(f"{tag}: vs device-arithmetic oracle: teacher rel-L2 {rt:.3e}  dense rel-L2 {rd:.3e}")
    assert rt <= 8e-3 and rd <= 8e-3
