"""Where does the on-device crop path spend its time?  CUDA-event time of the two batched calls vs host wall time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from types import SimpleNamespace
from clipself_b200.crops import _BatchCropper
cfg = SimpleNamespace(image_size=224)
dev = torch.device("cuda")
raw = bench.synth_raw_image_batch(cfg, 64, 32, 0)
cr = _BatchCropper()
for _ in range(2):
    cr(raw, dev)
torch.cuda.synchronize()
for it in range(3):
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); imgs, crops = cr(raw, dev); e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"device {e0.elapsed_time(e1):.2f} ms, host enqueue {1e3*(t1-t0):.2f} ms, wall {1e3*(t2-t0):.2f} ms; crops {tuple(crops.shape)} images {tuple(imgs.shape)}")
