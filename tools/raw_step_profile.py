"""Phase times of one cfg2 step fed by an image-backed batch (device crops) vs a host tensor batch: synchronised wall time
per phase (diagnostic only, not a benchmark)."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from clipself_b200 import ops, _lib
from clipself_b200.optim import FusedAdamW
from clipself_b200.training.clipself import CLIPSelf
dev = torch.device("cuda", 0)
wl = bench.WORKLOADS["cfg2"]
B, K = wl["batch"], wl["boxes"]
student, teacher = bench.build_models(wl["model"], dev)
cfg = student.visual.cfg
host = bench.synth_host_batch(cfg, B, K, wl["kind"], seed=1)
raw = bench.synth_raw_image_batch(cfg, B, K, seed=2)
margs = types.SimpleNamespace(multiscale=False, extract_type="v2", cosine_weight=1.0)
method = CLIPSelf()
opt = None
def step(batch, tag=None):
    global opt
    ts = [time.perf_counter()]
    def mark():
        if tag is not None:
            torch.cuda.synchronize()
        ts.append(time.perf_counter())
    losses, bs, _ = method(batch, student, teacher, None, dev, None, False, margs)
    mark()
    loss = losses["loss_cosine"]
    loss.backward()
    mark()
    if opt is None:
        opt = FusedAdamW(student.visual._student, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
    opt.step()
    v = loss.item()
    mark()
    if tag is not None:
        print(f"{tag}: forward(+inputs) {1e3*(ts[1]-ts[0]):.1f} ms, backward {1e3*(ts[2]-ts[1]):.1f} ms, opt+read {1e3*(ts[3]-ts[2]):.1f} ms, "
              f"total {1e3*(ts[3]-ts[0]):.1f} ms, encodes so far {_lib.lib().cs_tensor_map_encodes()}")
    return v
devb = tuple(t.to(dev) for t in host)
order = ((devb, "device"), (host, "host"), (raw, "raw")) if os.environ.get("WITH_DEV") else ((host, "host"), (raw, "raw"))
for b, name in order:
    for _ in range(4):
        step(b)
    torch.cuda.synchronize()
    for i in range(3):
        step(b, f"{name} step {i}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(5):
        step(b)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {1e3*(time.perf_counter()-t0)/5:.1f} ms per step wall, {e0.elapsed_time(e1)/5:.1f} ms between events (5 unsynchronised steps)")
if os.environ.get("PROFILE_RAW"):
    torch.cuda.cudart().cudaProfilerStart()
    step(raw)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
