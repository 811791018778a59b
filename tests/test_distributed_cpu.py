"""world_size-2 gloo test of the data-parallel contract: after the step's single
mean all-reduce over the flat gradient buffer every rank holds identical gradients equal to the mean
of the per-rank gradients — the test the reference's bypassed DDP (fact 7) would fail."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from clipself_b200.student import FlatLayout
    from clipself_b200.tower import TowerCfg
    cfg = TowerCfg(image_size=64, patch=16, width=128, heads=2, layers=3, hidden=384, embed_dim=64)
    lay = FlatLayout(cfg)
    g = torch.Generator().manual_seed(100 + rank)
    flat_grad = torch.randn(lay.n_total, generator=g)
    local = flat_grad.clone()
    # the ONE collective of the step (the function the autograd node calls), on the with-gradient prefix only
    from clipself_b200.student import allreduce_flat_gradient
    allreduce_flat_gradient(flat_grad, lay)
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    mean = torch.stack(gathered).mean(0)
    ok = torch.allclose(flat_grad[:lay.n_grad], mean[:lay.n_grad], atol=1e-6) and \
        torch.equal(flat_grad[lay.n_grad:], local[lay.n_grad:])            # grad-less tail untouched
    # partially unlocked tower (blocks[-2:] trainable): block 0's ranges must not be sent either
    part = local.clone()
    allreduce_flat_gradient(part, lay, first_trainable=1)
    ds, ns = lay.decay_start(1), lay.nodecay_start(1)
    ok = ok and torch.equal(part[:ds], local[:ds]) and torch.allclose(part[ds:lay.n_decay], mean[ds:lay.n_decay], atol=1e-6) \
        and torch.allclose(part[ns:lay.n_grad], mean[ns:lay.n_grad], atol=1e-6) and torch.equal(part[lay.n_grad:], local[lay.n_grad:])
    # rank-sharded synthetic batches: different seeds per rank, same shapes (weak scaling)
    from clipself_b200.data import synthetic_batch
    b = synthetic_batch(64, 2, 4, "grid", seed=1234 + rank)
    sums = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(sums, b[0].sum().reshape(1))
    ok = ok and len({round(float(s), 4) for s in sums}) == world
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_gradient_mean_allreduce_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
