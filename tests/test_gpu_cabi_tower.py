"""Tower-level C ABI (include/clipself_b200.h: cs_pack_weights_* / cs_query_workspace / cs_vit_forward_cls / _dense) called
through ctypes exactly as INTEGRATION.md shows a reference maintainer would, against the reference-generated fixtures."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import clipself_oracle as O

pytestmark = pytest.mark.gpu

CASES = {"tiny_grid": (O.CFG_TINY, 2, 4, "grid", False), "cfg1_b16": (O.CFG_B16, 2, 8, "grid", False),
         "l14_fwd": (O.CFG_L14_336, 1, 2, "proposal", False)}


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def _create(lib, L, ocfg, sd, dev):
    cfg = L.TowerCfgC(ocfg.image_size, ocfg.patch, ocfg.width, ocfg.heads, ocfg.layers, ocfg.hidden, ocfg.embed_dim, ocfg.pt_seq_len,
                      ocfg.ln_eps)
    need = C.c_int64(0)
    assert lib.cs_pack_weights_bytes(C.byref(cfg), C.byref(need)) == 0
    pack = torch.empty(need.value, dtype=torch.uint8, device=dev)
    tensors = {k: v.to(dev).float().contiguous() for k, v in sd.items()}
    names = (C.c_char_p * len(tensors))(*[k.encode() for k in tensors])
    ptrs = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors.values()])
    handle = C.c_void_p()
    rc = lib.cs_pack_weights_create(C.byref(cfg), names, ptrs, len(tensors), pack.data_ptr(), need.value, None, C.byref(handle))
    assert rc == 0, lib.cs_last_error().decode()
    torch.cuda.synchronize()
    return cfg, handle, pack


@pytest.mark.parametrize("tag", ["tiny_grid", "cfg1_b16", "l14_fwd"])
def test_vit_forward_through_the_c_abi(golden, tag):
    from clipself_b200 import _lib as L
    lib = L.lib()
    L.require_device()
    ocfg, B, K, kind, ragged = CASES[tag]
    g = golden(tag)
    seed = int(g["seed"])
    dev = torch.device("cuda")
    images, boxes, crops = O.synth_batch(ocfg, B, K, seed + 2, kind=kind, ragged=ragged)
    _, idx = O.extract_rois(boxes)
    tc = crops.flatten(0, 1)[idx].to(dev).contiguous()
    cfg, teacher, tpack = _create(lib, L, ocfg, O.synth_tower_weights(ocfg, seed + 1), dev)
    cfg2, student, spack = _create(lib, L, ocfg, O.synth_tower_weights(ocfg, seed), dev)
    n = tc.shape[0]
    chunk = max(n // 2, 1)                                   # several chunks through one workspace
    need = C.c_int64(0)
    assert lib.cs_query_workspace(C.byref(cfg), chunk, 0, C.byref(need)) == 0
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    out = torch.full((n, ocfg.embed_dim), float("nan"), device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    enc0 = lib.cs_tensor_map_encodes()
    results = []
    for it in range(4):                                      # eager, capture, replay, replay
        out.fill_(float("nan"))
        rc = lib.cs_vit_forward_cls(teacher, tc.data_ptr(), L.CS_F32, n, ws.data_ptr(), need.value, chunk, out.data_ptr(), stream)
        assert rc == 0, lib.cs_last_error().decode()
        torch.cuda.synchronize()
        results.append(out.clone())
        if it == 1:
            enc1 = lib.cs_tensor_map_encodes()
    assert lib.cs_tensor_map_encodes() == enc1, "steady-state calls must not encode TMA descriptors"
    assert all(torch.equal(results[0], r) for r in results[1:]), "graph replay must reproduce the eager launch bit for bit"
    rt = _rel(results[0].cpu().numpy(), g["teacher"])
    # dense map of the student tower
    imgs = images.to(dev).contiguous()
    dense = torch.empty(B, ocfg.grid, ocfg.grid, ocfg.embed_dim, device=dev)
    rc = lib.cs_vit_forward_dense(student, imgs.data_ptr(), L.CS_F32, B, 0, None, ws.data_ptr(), need.value, chunk, dense.data_ptr(), stream)
    assert rc == 0, lib.cs_last_error().decode()
    torch.cuda.synchronize()
    rd = _rel(dense.cpu().numpy(), g["dense_nhwc"])
    yard = float(g["ref_autocast_bf16_dense_rel_l2"])
    print(f"{tag}: C ABI teacher rel-L2 {rt:.3e}, dense rel-L2 {rd:.3e} (reference bf16-autocast yardstick {yard:.3e}); "
          f"{enc1 - enc0} descriptor encodes in the first two calls, 0 afterwards")
    assert rd <= 1.5 * yard + 2e-3 and rt <= 2.5e-2
    # the Python-sequenced path runs the identical kernels; its packer sums c1 / c2 in another order (1e-7), which the bf16
    # rounding points amplify to a few 1e-3 end to end (tests/test_gpu_parity_stages.py header) — same bound as there
    from clipself_b200.tower import TowerCfg, TowerEngine
    pcfg = TowerCfg(image_size=ocfg.image_size, patch=ocfg.patch, width=ocfg.width, heads=ocfg.heads, layers=ocfg.layers,
                    hidden=ocfg.hidden, embed_dim=ocfg.embed_dim, pt_seq_len=ocfg.pt_seq_len, ln_eps=ocfg.ln_eps)
    eng = TowerEngine(pcfg, O.synth_tower_weights(ocfg, seed + 1), dev)
    eng.native = None                                        # force the Python sequencing
    ref = eng.forward_cls(tc)
    assert _rel(results[0].cpu().numpy(), ref.cpu().numpy()) <= 8e-3
    # error behaviour: too small a workspace is refused with a message, nothing is launched
    rc = lib.cs_vit_forward_cls(teacher, tc.data_ptr(), L.CS_F32, n, ws.data_ptr(), 1024, chunk, out.data_ptr(), stream)
    assert rc != 0 and b"workspace too small" in lib.cs_last_error()
    assert lib.cs_pack_weights_destroy(teacher) == 0 and lib.cs_pack_weights_destroy(student) == 0


@pytest.mark.parametrize("tag", ["tiny_grid", "cfg1_b16"])
def test_cls_only_last_block_equals_the_full_block(tag, monkeypatch):
    """cs_vit_forward_cls runs the last block on the CLS rows only (encode_image reads norm(x)[:, 0], eva_vit_model.py:565-569):
    the result must be the one of the full block (CLIPSELF_FULL_LAST_BLOCK=1) up to the summation order of the CLS-query
    attention kernel."""
    from clipself_b200 import _lib as L
    lib = L.lib()
    L.require_device()
    ocfg, B, K, kind, ragged = CASES[tag]
    dev = torch.device("cuda")
    _, boxes, crops = O.synth_batch(ocfg, B, K, 77, kind=kind, ragged=ragged)
    _, idx = O.extract_rois(boxes)
    tc = crops.flatten(0, 1)[idx].to(dev).contiguous()
    n = tc.shape[0]
    sd = O.synth_tower_weights(ocfg, 78)
    outs = []
    for full in ("1", "0"):
        monkeypatch.setenv("CLIPSELF_FULL_LAST_BLOCK", full)            # read when the handle is created
        cfg, tower, pack = _create(lib, L, ocfg, sd, dev)
        need = C.c_int64(0)
        assert lib.cs_query_workspace(C.byref(cfg), n, 0, C.byref(need)) == 0
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        out = torch.full((n, ocfg.embed_dim), float("nan"), device=dev)
        rc = lib.cs_vit_forward_cls(tower, tc.data_ptr(), L.CS_F32, n, ws.data_ptr(), need.value, n, out.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.cs_last_error().decode()
        torch.cuda.synchronize()
        outs.append(out.clone())
        assert lib.cs_pack_weights_destroy(tower) == 0
    r = _rel(outs[1].cpu().numpy(), outs[0].cpu().numpy())
    print(f"{tag}: CLS-only last block vs full last block rel-L2 {r:.3e}")
    assert torch.isfinite(outs[1]).all() and r <= 2e-3


def test_pack_rejects_an_incomplete_state_dict():
    from clipself_b200 import _lib as L
    lib = L.lib()
    dev = torch.device("cuda")
    sd = O.synth_tower_weights(O.CFG_TINY, 3)
    sd.pop("blocks.1.mlp.w3.bias")
    cfg = L.TowerCfgC(64, 16, 128, 2, 3, 384, 64, 16, 1e-6)
    need = C.c_int64(0)
    lib.cs_pack_weights_bytes(C.byref(cfg), C.byref(need))
    pack = torch.empty(need.value, dtype=torch.uint8, device=dev)
    tensors = {k: v.to(dev).float().contiguous() for k, v in sd.items()}
    names = (C.c_char_p * len(tensors))(*[k.encode() for k in tensors])
    ptrs = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors.values()])
    handle = C.c_void_p()
    rc = lib.cs_pack_weights_create(C.byref(cfg), names, ptrs, len(tensors), pack.data_ptr(), need.value, None, C.byref(handle))
    assert rc != 0 and b"blocks.1.mlp.w3.bias" in lib.cs_last_error()
