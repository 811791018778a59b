"""Host-side logic that needs no GPU: flat parameter layout, state_dict compatibility, factory errors,
synthetic data contract, plug-in index extraction."""
import numpy as np
import pytest
import torch

from oracle import clipself_oracle as O


def test_state_dict_keys_match_reference_layout():
    """visual.* keys / shapes equal SURVEY.md Appendix D.1 (286 visual keys for B/16)."""
    from clipself_b200.factory import create_model
    m = create_model("EVA02-CLIP-B-16", "eva", cache_dir="")
    sd = m.state_dict()
    vis = {k: v for k, v in sd.items() if k.startswith("visual.")}
    assert len(vis) == 286
    expect = dict(O.tower_param_shapes(O.CFG_B16))
    for name, shape in expect.items():
        assert tuple(vis["visual." + name].shape) == shape, name
    rope_keys = [k for k in vis if "freqs_" in k]
    assert len(rope_keys) == 2 * (1 + 12)                      # visual.rope + every blocks.N.attn.rope
    assert "logit_scale" in sd
    assert m.visual.image_size == 224 and m.visual.patch_embed.patch_size == (16, 16) and m.embed_dim == 512


def test_rope_buffers_bit_exact_vs_oracle():
    from clipself_b200.tower import rope_tables, rope_vectors
    for grid in (14, 24, 4):
        cos, sin = rope_tables(grid, 64, 16)
        ocos, osin = O.rope_tables(grid, 64, 16)
        assert torch.equal(cos, ocos) and torch.equal(sin, osin)
        pos, freq = rope_vectors(grid, 64, 16)
        ang = torch.cat([(pos[:, None] * freq[None]).repeat_interleave(2, -1)[:, None].expand(grid, grid, 32),
                         (pos[:, None] * freq[None]).repeat_interleave(2, -1)[None].expand(grid, grid, 32)], -1)
        assert torch.equal(ang.reshape(-1, 64).cos(), cos)


def test_flat_layout_groups():
    from clipself_b200.student import FlatLayout
    from clipself_b200.tower import TowerCfg
    cfg = TowerCfg()
    lay = FlatLayout(cfg)
    D, Hd, L = cfg.width, cfg.hidden, cfg.layers
    per_block_w = 4 * D * D + 3 * D * Hd
    assert lay.n_decay == L * per_block_w - 2 * D * D                       # last block q/k weights live in the tail
    assert lay.n_total == 85_131_264 or lay.n_total > 85_000_000            # ~85.1 M trainable block params (SURVEY B)
    assert set(lay.gradless) == {f"blocks.{L-1}.attn.q_proj.weight", f"blocks.{L-1}.attn.k_proj.weight",
                                 f"blocks.{L-1}.attn.q_bias"}
    # q|k|v and w1|w2 adjacency that the fused GEMMs rely on
    for i in range(L - 1):
        q, k, v = (lay.offset[f"blocks.{i}.attn.{n}_proj.weight"] for n in "qkv")
        assert k == q + D * D and v == k + D * D
        assert lay.offset[f"blocks.{i}.mlp.w2.weight"] == lay.offset[f"blocks.{i}.mlp.w1.weight"] + Hd * D
        assert lay.offset[f"blocks.{i}.mlp.w2.bias"] == lay.offset[f"blocks.{i}.mlp.w1.bias"] + Hd
    assert all(o % 4 == 0 for o in lay.offset.values())                     # 16-byte aligned views
    # weight-decay grouping equals the reference rule (main.py:199-213): ndim < 2 -> no decay
    for name, off in lay.offset.items():
        if name in lay.gradless:
            assert off >= lay.n_grad
        elif len(lay.shape[name]) == 2:
            assert off < lay.n_decay
        else:
            assert lay.n_decay <= off < lay.n_grad


def test_factory_errors_like_reference():
    from clipself_b200.factory import create_model
    with pytest.raises(RuntimeError):
        create_model("no-such-model", "eva")
    with pytest.raises(RuntimeError):
        create_model("EVA02-CLIP-B-16", "eva", cache_dir="/nonexistent/ckpt.pt")      # eva_clip/factory.py:290-295
    with pytest.raises(RuntimeError):
        create_model("EVA02-CLIP-B-16", "eva", cache_dir="", require_pretrained=True)


def test_checkpoint_round_trip(tmp_path):
    from clipself_b200.factory import create_model
    a = create_model("EVA02-CLIP-B-16", "eva", cache_dir="")
    path = tmp_path / "ck.pt"
    torch.save({"state_dict": {"module." + k: v for k, v in a.state_dict().items()}}, path)
    b = create_model("EVA02-CLIP-B-16", "eva", cache_dir=str(path))
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)


def test_lock_semantics():
    from clipself_b200.factory import create_model
    m = create_model("EVA02-CLIP-B-16", "eva", cache_dir="")
    m.lock_image_tower(unlocked_groups=12)
    names = {n for n, p in m.visual.named_parameters() if p.requires_grad}
    assert names and all(n.startswith("blocks.") for n in names)
    m.lock_image_tower(unlocked_groups=3)
    names = {n.split(".")[1] for n, p in m.visual.named_parameters() if p.requires_grad}
    assert names == {"9", "10", "11"}
    m.lock_image_tower(unlocked_groups=0)                   # blocks[-0:] == all blocks, as in the reference
    names = {n.split(".")[1] for n, p in m.visual.named_parameters() if p.requires_grad}
    assert len(names) == 12


def test_synthetic_dataset_contract():
    from clipself_b200.data import SyntheticDistillDataset, grid_box_templates, synthetic_batch
    assert torch.equal(grid_box_templates(3, 3), O.grid_box_templates(3, 3))      # fp32 linspace edges, bit-exact
    ds = SyntheticDistillDataset(64, 32, 6, kind="grid", length=4, seed=1)
    img, boxes, crops = ds[2]
    assert img.shape == (3, 64, 64) and boxes.shape == (6, 5) and crops.shape == (6, 3, 32, 32)
    assert set(boxes[:, 4].tolist()) <= {0.0, 1.0} and boxes[0, 4] == 1.0
    images, b, c = synthetic_batch(64, 3, 5, "proposal", seed=2, ragged=True)
    valid = b[..., 4] > 0.5
    assert (b[valid][:, 2] > b[valid][:, 0]).all() and (b[~valid] == 0).all()


def test_plugin_host_index_extraction_bit_exact():
    """The host branch of the plug-in (CPU batch) selects exactly what clipself.py:29-36 selects."""
    _, boxes, crops = O.synth_batch(O.CFG_TINY, 4, 6, 5, kind="proposal", ragged=True, crop_size=8)
    rois_ref, idx_ref = O.extract_rois(boxes)
    boxes32 = boxes.float()
    valid = boxes32[:, :, 4] > 0.5
    rois = boxes32[valid][:, :4]
    assert torch.equal(rois, torch.cat(rois_ref))
    assert valid.flatten().nonzero().flatten().tolist() == idx_ref.tolist()
    assert torch.equal(crops[valid], crops.flatten(0, 1)[idx_ref])


def test_fused_adamw_group_bookkeeping():
    from clipself_b200.optim import FusedAdamW

    class FakeLayout:
        n_decay, n_grad = 8, 12

    class FakeEngine:
        layout = FakeLayout()
        device = torch.device("cpu")

    opt = FusedAdamW(FakeEngine(), lr=1e-3, weight_decay=0.1)
    assert [g["weight_decay"] for g in opt.param_groups] == [0.0, 0.1]
    for g in opt.param_groups:
        g["lr"] = 5e-4                                       # what scheduler.py:4-6 does
    assert opt.exp_avg.numel() == 12 and opt.state_dict()["step"] == 0


def test_open_clip_namespace_and_cliploss():
    """`import open_clip` / `training.clipself` resolve through compat/ ; ClipLoss matches the formula."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "compat"))
    try:
        import open_clip
        from training.clipself import CLIPSelf
        assert callable(open_clip.create_model) and CLIPSelf is not None
        g = torch.Generator().manual_seed(0)
        img = torch.nn.functional.normalize(torch.randn(6, 16, generator=g), dim=-1)
        txt = torch.nn.functional.normalize(torch.randn(6, 16, generator=g), dim=-1)
        loss = open_clip.ClipLoss()(img, txt, torch.tensor(10.0))
        logits = 10.0 * img @ txt.T
        ref = (torch.nn.functional.cross_entropy(logits, torch.arange(6)) +
               torch.nn.functional.cross_entropy(logits.T, torch.arange(6))) / 2
        assert torch.allclose(loss, ref)
        assert set(open_clip.ClipLoss()(img, txt, torch.tensor(10.0), output_dict=True)) == {"contrastive_loss"}
    finally:
        sys.path.remove(os.path.join(root, "compat"))
        for m in [k for k in sys.modules if k == "open_clip" or k.startswith("training")]:
            sys.modules.pop(m, None)


def test_teacher_chunk_schedule_covers_every_crop_once():
    from clipself_b200.tower import chunk_schedule
    for rows, step in [(2048, 256), (256, 256), (300, 256), (600, 256), (5, 4), (9, 4), (1, 1), (0, 1), (513, 256)]:
        pieces = chunk_schedule(rows, step)
        assert sum(n for _, n in pieces) == rows
        assert all(0 < n <= step for _, n in pieces)
        assert [s for s, _ in pieces] == [sum(n for _, n in pieces[:i]) for i in range(len(pieces))]
    assert chunk_schedule(2048, 256)[:3] == [(0, 64), (64, 128), (192, 256)]      # short pieces first


def test_flat_layout_trainable_suffix_ranges():
    """lock_image_tower(unlocked_groups=n) trains blocks[-n:]: their parameters must be exactly the two
    suffix ranges the optimizer / all-reduce use."""
    from clipself_b200.student import FlatLayout
    from clipself_b200.tower import TowerCfg
    lay = FlatLayout(TowerCfg(image_size=64, patch=16, width=128, heads=2, layers=3, hidden=384, embed_dim=64))
    assert (lay.decay_start(0), lay.nodecay_start(0)) == (0, lay.n_decay)
    for k in range(3):
        ds, ns = lay.decay_start(k), lay.nodecay_start(k)
        for name, o in lay.offset.items():
            if name in lay.gradless:
                assert o >= lay.n_grad
                continue
            inside = ds <= o < lay.n_decay or ns <= o < lay.n_grad
            assert inside == (int(name.split(".")[1]) >= k), (k, name)


def test_zero_shot_macc_matches_hand_computation():
    """macc_with_is_thing (zero_shot.py:135-169): per-class mean of top-1 / top-5 hits, thing and stuff apart."""
    from clipself_b200.training.zero_shot import macc_with_is_thing
    # 6 boxes: classes 0,0,2 are 'thing', 1,1,3 'stuff'; columns = top-5 hit matrix (at most one 1 per row)
    correct = torch.tensor([[1, 0, 0, 0, 0], [0, 0, 1, 0, 0], [0, 0, 0, 0, 0],
                            [0, 1, 0, 0, 0], [1, 0, 0, 0, 0], [0, 0, 0, 0, 1]], dtype=torch.float32)
    labels = torch.tensor([0, 0, 2, 1, 1, 3])
    is_thing = torch.tensor([1.0, 1.0, 1.0, 0.0, 0.0, 0.0])
    r = macc_with_is_thing(correct, is_thing, labels, "rois")
    assert r["rois.thing.macc1"] == pytest.approx((0.5 + 0.0) / 2)          # class 0: 1/2, class 2: 0/1
    assert r["rois.thing.macc5"] == pytest.approx((1.0 + 0.0) / 2)
    assert r["rois.stuff.macc1"] == pytest.approx((0.5 + 0.0) / 2)          # class 1: 1/2, class 3: 0/1
    assert r["rois.stuff.macc5"] == pytest.approx((1.0 + 1.0) / 2)


def test_zero_shot_run_with_a_stub_model():
    """The loop's bookkeeping (valid filtering, image-major order, top-5 / similarity extraction) with a stub
    model whose features are the class embeddings of the ground-truth labels: every box must be a top-1 hit."""
    import types
    from torch.utils.data import DataLoader
    from clipself_b200.data import SyntheticEvalDataset
    from clipself_b200.training import zero_shot

    ds = SyntheticEvalDataset(image_size=32, crop_size=16, max_boxes=5, num_classes=7, embed_dim=12, length=6, seed=3)
    emb = torch.nn.functional.normalize(torch.from_numpy(ds.embeddings).float(), dim=-1)

    class Stub:
        def __init__(self):
            self.labels = None
            self.visual = self

        def encode_boxes_and_masks(self, images, rois, masks, normalize=True):
            assert [r.shape[0] for r in rois] == [m.shape[0] for m in masks]
            n = sum(r.shape[0] for r in rois)
            lab = self.pending[:n]
            return emb[lab], emb[(lab + 1) % 7]                  # RoI features right, mask-pooled features wrong

        def encode_image(self, crops, normalize=True):
            return emb[self.pending[:crops.shape[0]]]

    stub = Stub()
    loader = DataLoader(ds, batch_size=3)
    # labels in the order the loop visits them
    order = []
    for _, bboxes, _, _, _ in loader:
        for b in bboxes:
            order.append(b[b[:, 5] > 0.5, 4].long())
    per_batch = [torch.cat(order[i:i + 3]) for i in range(0, len(order), 3)]

    class Feed:
        def __iter__(self_inner):
            for batch, lab in zip(loader, per_batch):
                stub.pending = lab
                yield batch
        dataset = ds

    args = types.SimpleNamespace(device="cpu", distributed=False, image_ave_pool=False)
    c_rois, c_crops, c_mask, s_rois, s_crops, s_mask, sizes, is_thing, labels = zero_shot.run(stub, Feed(), args)
    n = sum(x.numel() for x in per_batch)
    assert c_rois.shape == (n, 5) and labels.tolist() == torch.cat(per_batch).tolist()
    assert c_rois[:, 0].sum() == n and c_crops[:, 0].sum() == n           # always the top-1 hit
    assert c_mask[:, 0].sum() == 0                                         # never top-1 for the shifted labels
    torch.testing.assert_close(s_rois, torch.ones(n), atol=1e-5, rtol=0)
    assert sizes.shape == (n,) and is_thing.shape == (n,)


def test_crop_descriptors_match_the_oracle_rounding():
    """Host half of the on-device crop generation: rectangles, resized sizes, padding and the scratch bounds must
    agree with the oracle's restatement of Image.crop / ResizeMaxSize / ResizeLongest for every box."""
    import math
    import numpy as np
    from clipself_b200.crops import crop_descriptors
    from oracle import crops_oracle as C
    rng = np.random.default_rng(11)
    boxes = np.concatenate([rng.random((200, 2)) * [600, 400], rng.random((200, 2)) * [600, 400]], 1)
    boxes = np.concatenate([np.minimum(boxes[:, :2], boxes[:, 2:]), np.maximum(boxes[:, :2], boxes[:, 2:]) + 1.3], 1)
    boxes[:5] = [[0.5, 1.5, 20.5, 30.5], [2.5, 3.5, 10.49, 12.51], [0, 0, 640, 480], [10, 10, 10.2, 300], [5, 5, 6, 6]]
    for size, center in ((224, True), (32, True), (1024, False)):
        descs, ksize_max, rows_max = crop_descriptors(boxes, size, center)
        for b, d in zip(boxes, descs):
            rect = C.crop_box_to_rect(b)
            assert tuple(d[:4]) == rect
            w, h = rect[2] - rect[0], rect[3] - rect[1]
            if w <= 0 or h <= 0:
                assert d[4] == 0 and d[5] == 0
                continue
            nh, nw = C.resized_size(h, w, size)
            assert (d[5], d[4]) == (nh, nw)
            assert (d[6], d[7]) == (((size - nw) // 2, (size - nh) // 2) if center else (0, 0))
            for n_in, n_out in ((w, nw), (h, nh)):
                if n_out > 0:
                    assert C.precompute_coeffs(n_in, 0.0, float(n_in), n_out)[2] <= ksize_max
            assert h <= rows_max


def test_barrier_protocols_of_the_staged_attention_kernels():
    """tools/mbar_sim.py replays the mbarrier traffic of attention_bwd_tc.cu / attention_tc_long.cu under random
    schedules with asynchronous TMA / tcgen05.commit completions: no deadlock, no buffer overwritten while in use."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("mbar_sim", os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "tools", "mbar_sim.py"))
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    for seed in range(6):
        for items, nblk in [(1, 1), (2, 3), (3, 10)]:
            sim.sim_bwd(seed, items, nblk, dkv=False)
            sim.sim_bwd(seed, items, nblk, dkv=True)
            sim.sim_fwd_long(seed, items, nblk)
    # the model notices a protocol slip: a barrier sized for 32 of the 33 arrivals of the dKV producer warp
    import types
    with pytest.raises(AssertionError):
        s = sim.Sim(0)
        s.bar("b", 32)
        s.bars["b"].arrive(33)


def test_staged_attention_backward_dataflow_reproduces_autograd():
    """tools/attn_bwd_emul.py replays the operand table / masks / accumulator-to-section mapping / inverse RoPE of
    csrc/attention_bwd_tc.cu in numpy (tensor-core products as matmuls) and compares with torch autograd."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("attn_bwd_emul", os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "tools", "attn_bwd_emul.py"))
    emul = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(emul)
    for case in [(2, 17, 2, True), (2, 64, 1, False), (1, 130, 1, False)]:
        res = emul.check(*case)
        assert all(v < 2e-2 for v in res.values()), (case, res)


def test_bench_flop_model_counts_the_cls_only_last_block():
    """bench.py reports the reference's algorithmic FLOPs (the contract's metric) and, separately, what the kernels execute:
    the teacher's last block runs on the CLS row only, which removes ~5.7 % of a cfg2 step's work."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from clipself_b200.tower import TowerCfg
    cfg = TowerCfg(image_size=224, patch=16, width=768, heads=12, layers=12, hidden=2048, embed_dim=512, pt_seq_len=16, ln_eps=1e-6)
    algo = bench.flops_per_image(cfg, 32)
    done = bench.flops_per_image(cfg, 32, executed=True)
    assert abs(algo / 1e9 - 1228.1) < 0.1
    assert 0.05 < 1.0 - done / algo < 0.065
    assert bench.flops_per_image(cfg, 0, executed=True) == bench.flops_per_image(cfg, 0)      # the student is not affected


def test_batch_cropper_buffers_keep_their_address():
    """The on-device cropper hands out views of persistent buffers: same shape -> same storage (the native tower's CUDA graphs
    and TMA descriptors are cached per address), a larger request re-allocates, a smaller one reuses."""
    import torch
    from clipself_b200.crops import _BatchCropper
    c = _BatchCropper()
    dev = torch.device("cpu")
    a = c._buffer(("crops", "out"), (4, 3, 8, 8), torch.float32, dev)
    b = c._buffer(("crops", "out"), (4, 3, 8, 8), torch.float32, dev)
    assert a.data_ptr() == b.data_ptr() and a.shape == (4, 3, 8, 8)
    small = c._buffer(("crops", "out"), (2, 3, 8, 8), torch.float32, dev)
    assert small.data_ptr() == a.data_ptr() and small.shape == (2, 3, 8, 8)
    big = c._buffer(("crops", "out"), (8, 3, 8, 8), torch.float32, dev)
    assert big.numel() == 8 * 3 * 64
    other = c._buffer(("det", "out"), (4, 3, 8, 8), torch.float32, dev)
    assert other.data_ptr() != big.data_ptr()
    x = torch.arange(6, dtype=torch.float32).view(2, 3)
    assert c.cast(x, torch.float32) is x
    y = c.cast(x, torch.bfloat16)
    assert y.dtype == torch.bfloat16 and torch.equal(y.float(), x) and c.cast(x, torch.bfloat16).data_ptr() == y.data_ptr()
