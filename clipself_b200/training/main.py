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


class LossScaler:
    """The `torch.cuda.amp.GradScaler` protocol the reference runs under `--precision amp` (main.py:214,
    train.py:46-50 `scaler.scale(loss).backward()`, :98-111 unscale / clip / `scaler.step` / `scaler.update`):
    dynamic loss scale (init 2^16, x2 after 2000 clean steps, x0.5 and the optimizer step SKIPPED when a
    gradient is inf/nan), `state_dict()` with torch's key names so `checkpoint['scaler']` round-trips
    (main.py:311-312, :231-232).  The unscale is folded into the fused AdamW's gradient factor.  The tensor-core
    operands stay bf16 (this library has no fp16 kernels); the protocol, the skipped steps and the checkpoint
    entry are the reference's."""

    def __init__(self, init_scale=2.0 ** 16, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.scale_value, self.growth_factor, self.backoff_factor = float(init_scale), growth_factor, backoff_factor
        self.growth_interval, self.growth_tracker, self.found_inf = growth_interval, 0, False

    def scale(self, loss):
        return loss * self.scale_value

    def check(self, flat_grad_span) -> bool:
        """True when every gradient is finite (one reduction + a host read, like GradScaler's found_inf)."""
        self.found_inf = not bool(torch.isfinite(flat_grad_span.sum()))
        return not self.found_inf

    def update(self):
        if self.found_inf:
            self.scale_value *= self.backoff_factor
            self.growth_tracker = 0
        else:
            self.growth_tracker += 1
            if self.growth_tracker == self.growth_interval:
                self.scale_value *= self.growth_factor
                self.growth_tracker = 0

    def state_dict(self):
        return {"scale": self.scale_value, "growth_factor": self.growth_factor, "backoff_factor": self.backoff_factor,
                "growth_interval": self.growth_interval, "_growth_tracker": self.growth_tracker}

    def load_state_dict(self, sd):
        self.scale_value, self.growth_factor = float(sd["scale"]), sd["growth_factor"]
        self.backoff_factor, self.growth_interval = sd["backoff_factor"], sd["growth_interval"]
        self.growth_tracker = int(sd["_growth_tracker"])


LATEST_CHECKPOINT_NAME = "epoch_latest.pt"        # main.py:40


def save_checkpoints(args, checkpoint_dict, completed_epoch, out_dir):
    """main.py:300-328: epoch_N.pt at the last epoch or every --save-frequency epochs, optional removal of the
    previous one, and --save-most-recent through tmp.pt + os.replace so a failed save cannot corrupt the latest."""
    os.makedirs(out_dir, exist_ok=True)
    if completed_epoch == args.epochs or (args.save_frequency > 0 and completed_epoch % args.save_frequency == 0):
        torch.save(checkpoint_dict, os.path.join(out_dir, f"epoch_{completed_epoch}.pt"))
    if args.delete_previous_checkpoint:
        previous = os.path.join(out_dir, f"epoch_{completed_epoch - 1}.pt")
        if os.path.exists(previous):
            os.remove(previous)
    if args.save_most_recent:
        tmp, latest = os.path.join(out_dir, "tmp.pt"), os.path.join(out_dir, LATEST_CHECKPOINT_NAME)
        torch.save(checkpoint_dict, tmp)
        os.replace(tmp, latest)


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
    if args.precision == "fp32":
        raise NotImplementedError("--precision fp32: this library has no fp32-operand path (every contraction runs on "
                                  "the tensor cores with bf16 operands, fp32 accumulation); use amp_bf16 or amp")
    # precision.py:5-12 / main.py:214: `amp` = autocast + GradScaler, `amp_bf16` = bf16 autocast without a scaler
    scaler = LossScaler() if args.precision == "amp" else None
    if scaler is not None and rank == 0:
        logging.warning("--precision amp: running the GradScaler protocol (dynamic loss scale, skipped steps on inf/nan, "
                        "'scaler' in checkpoints) over bf16 tensor-core operands; there are no fp16 kernels in this library")
    if args.accum_freq != 1:
        raise AssertionError("accum_freq must be 1 (train.py:89)")
    if args.dataset_type not in ("synthetic_distill", "synthetic_images_distill", "grid_distill"):
        raise NotImplementedError(f"--dataset-type {args.dataset_type}: only grid_distill (image files, crops made on the "
                                  "device) and the synthetic_* types are built in (SURVEY.md §2: the COCO proposal / "
                                  "RegionCLIP pipelines are out of scope)")

    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    # `--cache-dir` is the checkpoint path (scripts/*.sh); a path that does not exist raises inside create_model
    # exactly like eva_clip/factory.py:290-295.  Random initialisation is an explicit opt-in: --cache-dir "".
    cache = args.cache_dir or ""
    model = create_model(args.model, "eva", precision="amp_bf16", device=device, cache_dir=cache)
    dist_model = create_model(args.model, "eva", precision="amp_bf16", device=device, cache_dir=cache)
    if not cache:
        if rank == 0:
            logging.warning("--cache-dir '': RANDOM-INIT run; the teacher is a copy of the random student "
                            "(synthetic benchmarking only: nothing is being distilled)")
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
        model.load_state_dict(sd)                               # strict, like main.py:223-234
        if "epoch" in ckpt:
            start_epoch, resume_opt = int(ckpt["epoch"]), ckpt.get("optimizer")
            if scaler is not None and "scaler" in ckpt:
                scaler.load_state_dict(ckpt["scaler"])
        if rank == 0:
            logging.info(f"=> resuming checkpoint '{args.resume}' (epoch {start_epoch})")
    trainable_outside_blocks = [n for n, p in model.named_parameters()
                                if p.requires_grad and not n.startswith("visual.blocks.") and n != "logit_scale"]
    if trainable_outside_blocks:
        raise NotImplementedError(
            "the fused training path updates visual.blocks.* only (what the reference's scripts train with --lock-image); "
            f"these parameters would silently stay constant: {trainable_outside_blocks[:6]} — pass --lock-image")
    model.train()
    dist_model.eval()
    model.visual.overlap_gradient_sync = args.distributed         # all-reduce + AdamW under the next step's teacher
    method = CLIPSelf()

    # student images at --det-image-size (scripts: 1024 / 896), teacher crops at the tower's own size
    # (data.py:226-245: crops are resized to args.input_size)
    collate = None
    if args.dataset_type == "grid_distill":
        # the reference's GridDistillDataset (data.py:135-281) with the pixel work moved to the device: the workers decode
        # the files to uint8 and do the box arithmetic, the plug-in makes the student image and the K crops (crops.py)
        from ..data import ImageGridDistillDataset
        if not args.train_data:
            raise RuntimeError("--dataset-type grid_distill needs --train-data (COCO-style json or an image directory)")
        dataset = ImageGridDistillDataset(args.train_data, args.train_image_root, args.det_image_size, args.input_size,
                                          args.max_boxes, args.max_split, args.crop_scale, seed=args.seed)
        collate = dataset.collate
    elif args.dataset_type == "synthetic_images_distill":
        from ..data import SyntheticImageGridDataset
        dataset = SyntheticImageGridDataset(args.det_image_size, args.input_size, args.max_boxes, args.max_split, args.crop_scale,
                                            length=max(args.batch_size * world * 8, 64), seed=args.seed)
        collate = dataset.collate
    else:
        dataset = SyntheticDistillDataset(args.det_image_size, args.input_size, args.max_boxes,
                                          kind="grid", length=max(args.batch_size * world * 8, 64), seed=args.seed)
    sampler = DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=True, seed=args.seed) \
        if args.distributed else None
    loader = DataLoader(dataset, batch_size=args.batch_size, shuffle=sampler is None, sampler=sampler,
                        num_workers=args.workers, pin_memory=collate is None, drop_last=True, collate_fn=collate)
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
        t_last, i_last = time.time(), -1
        it = iter(loader)
        for i in range(steps_per_epoch):
            try:
                batch = next(it)
            except StopIteration:
                it = iter(loader)
                batch = next(it)
            losses, batch_size, logit_scale = method(batch, model, dist_model, None, device, None, args.distributed, args)
            total_loss = sum(losses.values())
            (scaler.scale(total_loss) if scaler is not None else total_loss).backward()     # train.py:46-50
            if optimizer is None:                                  # engine exists after the first forward
                optimizer = FusedAdamW(model.visual._student, lr=args.lr, betas=(args.beta1, args.beta2),
                                       eps=args.eps, weight_decay=args.wd)
                if resume_opt is not None:
                    optimizer.load_state_dict({k: (v.to(device) if torch.is_tensor(v) else v) for k, v in resume_opt.items()})
            if not args.skip_scheduler:
                for g in optimizer.param_groups:
                    g["lr"] = scheduler(step)
            eng = model.visual._student
            if args.grad_clip_norm is not None or scaler is not None:
                eng.wait_gradients()                               # the norm / finiteness checks read the reduced gradient
            span = eng.flat_grad[eng.layout.decay_start(eng.first_trainable):eng.layout.n_grad]
            grad_scale = 1.0 / scaler.scale_value if scaler is not None else 1.0     # scaler.unscale_ (train.py:108)
            finite = scaler.check(span) if scaler is not None else True
            if finite and args.grad_clip_norm is not None:         # train.py:107-114 (clip_grad_norm_, L2, unscaled grads)
                norm = float(torch.linalg.vector_norm(span)) * grad_scale
                grad_scale *= min(1.0, args.grad_clip_norm / (norm + 1e-6))
            if finite:
                optimizer.step(grad_scale=grad_scale)              # unscale and clip factors applied inside the fused update
            if scaler is not None:
                scaler.update()                                    # train.py:111
            with torch.no_grad():                                  # train.py:118-119
                model.logit_scale.clamp_(0, math.log(100))
            step += 1
            if rank == 0 and (i % args.log_every_n_steps == 0 or i == steps_per_epoch - 1):
                loss_v = total_loss.item()
                dt = time.time() - t_last
                n = i - i_last                                     # steps since the previous log line
                logging.info(f"Train Epoch: {epoch} [{i + 1}/{steps_per_epoch}] Loss: {loss_v:.5f} "
                             f"LR: {optimizer.param_groups[0]['lr']:.3e} Logit Scale: {logit_scale.item():.3f} "
                             f"{n * batch_size * world / max(dt, 1e-9):.1f} samples/s")
                t_last, i_last = time.time(), i
        if rank == 0 and args.name:                                # main.py:280-328
            sd = model.state_dict()
            if args.alpha < 1.0:
                sd = student_teacher_ensemble(sd, dist_model.state_dict(), args.alpha)
            checkpoint_dict = {"epoch": epoch + 1, "name": args.name, "state_dict": {k: v.cpu() for k, v in sd.items()},
                               "optimizer": {k: (v.cpu() if torch.is_tensor(v) else v)
                                             for k, v in optimizer.state_dict().items()}}
            if scaler is not None:
                checkpoint_dict["scaler"] = scaler.state_dict()
            save_checkpoints(args, checkpoint_dict, epoch + 1, os.path.join(args.logs, args.name, "checkpoints"))
        if val_loader is not None:                                 # main.py:330 -> zero_shot.py:172-193 (frequency checked there)
            model.eval()
            metrics = zero_shot_eval(model, {"val": val_loader}, epoch + 1, args)
            model.train()
            if rank == 0 and metrics:
                logging.info(f"Eval Epoch: {epoch + 1} " + ", ".join(f"{k}: {v:.4f}" for k, v in metrics.items()))
    if args.distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
