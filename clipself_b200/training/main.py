"""`python -m clipself_b200.training.main` — the training driver around the hot path.

Mirrors the step loop of src/training/train.py:62-165 and the set-up of src/training/main.py:55-342
for the CLIPSelf method: model + teacher creation, lock, AdamW with the reference's two parameter
groups, cosine schedule with warm-up applied before each step (scheduler.py:43-53, train.py:84-85),
logit_scale clamp (train.py:118-119), samples/s logging (train.py:143-151) and the epoch-end
student/teacher weight ensemble checkpoint (train.py:53-59, main.py:280-317).

Datasets: only the additive `--dataset-type synthetic_distill` is built in (the COCO/PIL pipelines
are out of scope); everything flows through the same plug-in boundary.
"""
from __future__ import annotations

import logging
import math
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
from torch.utils.data import DataLoader
from torch.utils.data.distributed import DistributedSampler

from ..data import SyntheticDistillDataset
from ..factory import create_model
from ..optim import FusedAdamW
from .zero_shot import zero_shot_eval
from .clipself import CLIPSelf
from .params import parse_args


def cosine_lr(base_lr, warmup_length, steps):
    """scheduler.py:43-53 (value applied BEFORE the step)."""
    def lr_at(step):
        if step < warmup_length:
            return base_lr * (step + 1) / warmup_length
        e, es = step - warmup_length, steps - warmup_length
        return 0.5 * (1 + math.cos(math.pi * e / es)) * base_lr
    return lr_at


def student_teacher_ensemble(student_sd, teacher_sd, alpha):
    """train.py:53-59: alpha * student + (1 - alpha) * teacher, key by key."""
    return {k: (v * alpha + teacher_sd[k].to(v.device) * (1.0 - alpha)) if v.is_floating_point() else v
            for k, v in student_sd.items()}


def main(argv=None):
    args = parse_args(argv)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.distributed = world > 1
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if args.distributed:
        dist.init_process_group(args.dist_backend, init_method=args.dist_url, device_id=device)
    logging.basicConfig(level=logging.INFO if rank == 0 else logging.WARN, stream=sys.stdout,
                        format="%(asctime)s | %(levelname)s | %(message)s")
    if args.precision in ("bf16", "fp16"):
        raise NotImplementedError("pure bf16/fp16 cannot run the reference path either (SURVEY fact 8): use amp_bf16")
    if args.accum_freq != 1:
        raise AssertionError("accum_freq must be 1 (train.py:89)")
    if args.dataset_type != "synthetic_distill":
        raise NotImplementedError(f"--dataset-type {args.dataset_type}: the COCO/PIL data pipeline is out of scope "
                                  "(SURVEY.md §2); use synthetic_distill")

    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    cache = args.cache_dir if args.cache_dir and os.path.exists(args.cache_dir) else ""
    model = create_model(args.model, "eva", precision="amp_bf16", device=device, cache_dir=cache)
    dist_model = create_model(args.model, "eva", precision="amp_bf16", device=device, cache_dir=cache)
    if not cache:
        dist_model.load_state_dict(model.state_dict())          # random init: teacher = student copy
    args.input_size = model.visual.image_size
    if args.lock_image:
        model.lock_image_tower(unlocked_groups=args.lock_image_unlocked_groups)
    # optionally resume (main.py:216-235): a train checkpoint {epoch, state_dict, optimizer} continues at its epoch
    # (the optimizer moments are restored when the fused optimizer is created after the first forward); a bare
    # state_dict is loaded for fine-tuning / evaluation.  `text.*` keys are not on this path and are skipped.
    start_epoch, resume_opt = 0, None
    if args.resume:
        ckpt = torch.load(args.resume, map_location="cpu")
        sd = ckpt["state_dict"] if "epoch" in ckpt else ckpt
        if next(iter(sd)).startswith("module"):
            sd = {k[len("module."):]: v for k, v in sd.items()}
        missing, unexpected = model.load_state_dict(sd, strict=False)
        bad = [k for k in list(missing) + list(unexpected) if not (k.startswith("text.") or "rope" in k)]
        if bad:
            raise RuntimeError(f"--resume {args.resume}: state_dict mismatch on {bad[:8]}")
        if "epoch" in ckpt:
            start_epoch, resume_opt = int(ckpt["epoch"]), ckpt.get("optimizer")
        if rank == 0:
            logging.info(f"=> resuming checkpoint '{args.resume}' (epoch {start_epoch})")
    model.train()
    dist_model.eval()
    method = CLIPSelf()

    # student images at --det-image-size (scripts: 1024 / 896), teacher crops at the tower's own size
    # (data.py:226-245: crops are resized to args.input_size)
    dataset = SyntheticDistillDataset(args.det_image_size, args.input_size, args.max_boxes,
                                      kind="grid", length=max(args.batch_size * world * 8, 64), seed=args.seed)
    sampler = DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=True, seed=args.seed) \
        if args.distributed else None
    loader = DataLoader(dataset, batch_size=args.batch_size, shuffle=sampler is None, sampler=sampler,
                        num_workers=args.workers, pin_memory=True, drop_last=True)
    val_loader = None
    if args.synthetic_eval_classes > 0:
        from ..data import SyntheticEvalDataset
        args.device = device
        val_set = SyntheticEvalDataset(args.det_image_size, args.input_size, args.max_boxes, args.synthetic_eval_classes,
                                       model.embed_dim, args.downsample_factor, length=max(args.batch_size * 2, 8),
                                       seed=args.seed + 17)
        val_loader = DataLoader(val_set, batch_size=args.batch_size, shuffle=False, num_workers=0, pin_memory=True)
    steps_per_epoch = args.train_steps_per_epoch or len(loader)
    total_steps = steps_per_epoch * args.epochs
    scheduler = cosine_lr(args.lr, args.warmup, total_steps)

    optimizer = None
    step = start_epoch * steps_per_epoch
    for epoch in range(start_epoch, args.epochs):
        if sampler is not None:
            sampler.set_epoch(epoch)
        t_last = time.time()
        it = iter(loader)
        for i in range(steps_per_epoch):
            try:
                batch = next(it)
            except StopIteration:
                it = iter(loader)
                batch = next(it)
            losses, batch_size, logit_scale = method(batch, model, dist_model, None, device, None, args.distributed, args)
            total_loss = sum(losses.values())
            total_loss.backward()
            if optimizer is None:                                  # engine exists after the first forward
                optimizer = FusedAdamW(model.visual._student, lr=args.lr, betas=(args.beta1, args.beta2),
                                       eps=args.eps, weight_decay=args.wd)
                if resume_opt is not None:
                    optimizer.load_state_dict({k: (v.to(device) if torch.is_tensor(v) else v) for k, v in resume_opt.items()})
            if not args.skip_scheduler:
                for g in optimizer.param_groups:
                    g["lr"] = scheduler(step)
            grad_scale = 1.0
            if args.grad_clip_norm is not None:                    # train.py:107-114 (clip_grad_norm_, L2)
                eng = model.visual._student
                span = eng.flat_grad[eng.layout.decay_start(eng.first_trainable):eng.layout.n_grad]
                grad_scale = min(1.0, args.grad_clip_norm / (float(torch.linalg.vector_norm(span)) + 1e-6))
            optimizer.step(grad_scale=grad_scale)                  # the clip factor is applied inside the fused update
            with torch.no_grad():                                  # train.py:118-119
                model.logit_scale.clamp_(0, math.log(100))
            step += 1
            if rank == 0 and (i % args.log_every_n_steps == 0 or i == steps_per_epoch - 1):
                loss_v = total_loss.item()
                dt = time.time() - t_last
                n = min(args.log_every_n_steps, i + 1) if i else 1
                logging.info(f"Train Epoch: {epoch} [{i + 1}/{steps_per_epoch}] Loss: {loss_v:.5f} "
                             f"LR: {optimizer.param_groups[0]['lr']:.3e} Logit Scale: {logit_scale.item():.3f} "
                             f"{n * batch_size * world / max(dt, 1e-9):.1f} samples/s")
                t_last = time.time()
        if rank == 0 and args.name and (epoch + 1) % args.save_frequency == 0:
            out_dir = os.path.join(args.logs, args.name, "checkpoints")
            os.makedirs(out_dir, exist_ok=True)
            sd = model.state_dict()
            if args.alpha < 1.0:
                sd = student_teacher_ensemble(sd, dist_model.state_dict(), args.alpha)
            torch.save({"epoch": epoch + 1, "name": args.name, "state_dict": {k: v.cpu() for k, v in sd.items()},
                        "optimizer": {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in optimizer.state_dict().items()}},
                       os.path.join(out_dir, f"epoch_{epoch + 1}.pt"))
        if val_loader is not None:                                 # train.py:168-187 -> zero_shot.py:172-193
            model.eval()
            metrics = zero_shot_eval(model, {"val": val_loader}, epoch + 1, args)
            model.train()
            if rank == 0 and metrics:
                logging.info(f"Eval Epoch: {epoch + 1} " + ", ".join(f"{k}: {v:.4f}" for k, v in metrics.items()))
    if args.distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
