"""CUDA-event timing of the region-path kernels at BASELINE cfg2 / cfg5 sizes: achieved HBM GB/s against the
algorithmic bytes of SURVEY.md §8d (L2 flushed between launches)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clipself_b200 import ops
from clipself_b200.data import synthetic_batch
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0


def timeit(fn, reps=9):
    """Device time of one invocation: the launches are captured in a CUDA graph so that the host-side
    ctypes/allocator overhead (10-20 us per call, longer than most of these kernels) is not measured."""
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3      # us


def report(name, us, nbytes):
    gbs = nbytes / us / 1e3
    print(f"{name:44s} {us:8.1f} us  {nbytes/1e6:8.2f} MB  {gbs:8.1f} GB/s  {100*gbs/peak:5.1f} % of {peak:.0f} GB/s (measured copy peak)")


for tag, (B, K, g, C, S) in {"cfg2 (B/16: 64 img x 32 boxes, 14x14x512)": (64, 32, 14, 512, 224),
                              "cfg5 (L/14: 32 img x 64 boxes, 24x24x768)": (32, 64, 24, 768, 336)}.items():
    print("---", tag)
    _, boxes, _ = synthetic_batch(S, B, K, "grid", 1, crop_size=8)
    boxes = boxes.to(dev)
    fmap = torch.nn.functional.normalize(torch.randn(B, g, g, C, device=dev), dim=-1)
    R = B * K
    rois, crop_index, roi_batch, offsets = ops.extract_rois(boxes)
    report("extract_rois", timeit(lambda: ops.extract_rois(boxes)), B * K * 20 + R * 24 + (B + 1) * 4)
    out, wy, wx = ops.roi_align_fwd(fmap, rois, offsets, R)
    report("roi_align fwd (weights + pool)", timeit(lambda: ops.roi_align_fwd(fmap, rois, offsets, R)),
           fmap.numel() * 4 + R * 16 + R * C * 4)
    d_out = torch.randn(R, C, device=dev)
    report("roi_align bwd (gather form)", timeit(lambda: ops.roi_align_bwd(d_out, fmap.shape, offsets, R, wy, wx)),
           R * C * 4 + fmap.numel() * 4)
    t = torch.randn(R, C, device=dev)
    loss, stats = ops.cosine_loss_fwd(out, t, 1.0)
    report("normalise + cosine loss fwd", timeit(lambda: ops.cosine_loss_fwd(out, t, 1.0)), 2 * R * C * 4 + R * 12 + 4)
    one = torch.ones((), device=dev)
    report("normalise + cosine loss bwd", timeit(lambda: ops.cosine_loss_bwd(out, t, stats, 1.0, one)), 3 * R * C * 4)
    masks = (torch.rand(R, g * g, device=dev) > 0.5).float()
    report("mask_pool fwd", timeit(lambda: ops.mask_pool_fwd(fmap.view(B, g * g, C), masks, offsets)),
           fmap.numel() * 4 + masks.numel() * 4 + R * C * 4)
    x = torch.randn(B * g * g, C, device=dev)
    y, inv = ops.l2norm_fwd(x)
    report("per-token l2norm fwd", timeit(lambda: ops.l2norm_fwd(x)), 2 * x.numel() * 4)
