"""Two-GPU NCCL test of the data-parallel contract (the test the reference's bypassed DDP would fail):
after ONE real distillation step with `distributed=True` the student gradients are bit-identical on both ranks and equal
the mean of the gradients each rank computes alone on its own shard.  Needs 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import sys
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clipself_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from clipself_b200.model import CustomCLIP
    from clipself_b200.training.clipself import CLIPSelf
    ocfg = O.CFG_TINY

    def build_model(ocfg, seed, dev):
        vis = dict(image_size=ocfg.image_size, layers=ocfg.layers, width=ocfg.width, head_width=64,
                   patch_size=ocfg.patch, mlp_ratio=ocfg.hidden / ocfg.width, pt_hw_seq_len=ocfg.pt_seq_len)
        m = CustomCLIP(embed_dim=ocfg.embed_dim, vision_cfg=vis)
        m.visual.load_state_dict(O.synth_tower_weights(ocfg, seed), strict=False)
        return m.to(dev)

    student, teacher = build_model(ocfg, 41, dev), build_model(ocfg, 42, dev)          # same weights on both ranks
    student.lock_image_tower(unlocked_groups=ocfg.layers)
    student.train()
    teacher.eval()
    args = types.SimpleNamespace(multiscale=False, extract_type="v2", cosine_weight=1.0)
    method = CLIPSelf()

    def grads_of(batch_rank, distributed):
        batch = O.synth_batch(ocfg, 3, 5, 50 + batch_rank, kind="proposal", ragged=True)     # a different shard per rank
        losses, _, _ = method(batch, student, teacher, None, dev, None, distributed, args)
        losses["loss_cosine"].backward()
        torch.cuda.synchronize()
        eng = student.visual._student
        return eng.flat_grad[:eng.layout.n_grad].clone()

    local = [grads_of(r, False) for r in range(world)]              # every rank's shard, no collective
    synced = grads_of(rank, True)                                   # the real step: one NCCL mean all-reduce
    mean = torch.stack(local).mean(0)
    gathered = [torch.empty_like(synced) for _ in range(world)]
    dist.all_gather(gathered, synced)
    identical = all(torch.equal(gathered[0], g) for g in gathered)
    err = ((synced - mean).norm() / mean.norm()).item()
    differ = ((local[0] - local[1]).norm() / local[0].norm()).item()
    # the overlapped form (side-stream all-reduce + fused AdamW, teacher first): same gradients, and after the optimizer
    # step the weights are bit-identical on every rank
    from clipself_b200.optim import FusedAdamW
    student.visual.overlap_gradient_sync = True
    synced2 = grads_of(rank, True)
    eng = student.visual._student
    eng.wait_gradients()
    torch.cuda.synchronize()
    same_as_blocking = torch.equal(synced2[: eng.layout.n_grad], eng.flat_grad[: eng.layout.n_grad]) and torch.equal(synced2, synced)
    opt = FusedAdamW(eng, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
    opt.step()
    assert eng.weights_ready is not None
    eng.wait_weights()
    torch.cuda.synchronize()
    w = eng.flat_param.clone()
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    weights_identical = all(torch.equal(ws[0], x) for x in ws)
    out[rank] = (bool(identical), err, differ, bool(same_as_blocking), bool(weights_identical))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_one_step_gradients_identical_across_ranks_and_equal_the_mean():
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29500 + os.getpid() % 2000, out), nprocs=world, join=True)
    for rank in range(world):
        identical, err, differ, same_as_blocking, weights_identical = out[rank]
        print(f"rank {rank}: identical across ranks {identical}, |synced - mean| / |mean| = {err:.2e}, shards differ by {differ:.2f}")
        assert identical
        assert err <= 1e-6            # NCCL AVG over 2 ranks in f32 vs torch mean: rounding of one add + one scale
        assert differ > 1e-2          # the two shards really had different gradients
        assert same_as_blocking and weights_identical
