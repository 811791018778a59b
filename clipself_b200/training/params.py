"""Command line of `python -m training.main` (src/training/params.py:26-476) — the flags the reference's
launch scripts pass (scripts/train_clipself_coco_image_patches_eva_vit*.sh) plus the additive
`--dataset-type synthetic_distill` of SURVEY.md §7.  Defaults equal the reference's."""
import argparse


def get_default_params(model_name):
    """params.py:5-11: `"vit" in "eva02-clip-b-16"` is False -> the non-ViT defaults."""
    model_name = model_name.lower()
    if "vit" in model_name:
        return {"lr": 5.0e-4, "beta1": 0.9, "beta2": 0.98, "eps": 1.0e-6}
    return {"lr": 5.0e-4, "beta1": 0.9, "beta2": 0.999, "eps": 1.0e-8}


def parse_args(args=None):
    p = argparse.ArgumentParser()
    p.add_argument("--max-boxes", type=int, default=20)
    p.add_argument("--min-size", type=int, default=8)
    p.add_argument("--max-size", type=int, default=1024)
    p.add_argument("--train-data", type=str, default=None)
    p.add_argument("--val-data", type=str, default=None)
    p.add_argument("--embed-path", type=str, default=None)
    p.add_argument("--train-image-root", type=str, default=None)
    p.add_argument("--val-image-root", type=str, default=None)
    p.add_argument("--dataset-type", default="grid_distill",
                   choices=["proposals_distill", "region_clip", "grid_distill", "synthetic_distill", "synthetic_images_distill"])
    p.add_argument("--test-type", default="coco_panoptic")
    p.add_argument("--max-split", type=int, default=6)
    p.add_argument("--logs", type=str, default="./logs/")
    p.add_argument("--name", type=str, default=None)
    p.add_argument("--workers", type=int, default=1)
    p.add_argument("--batch-size", type=int, default=64)
    p.add_argument("--epochs", type=int, default=32)
    p.add_argument("--lr", type=float, default=None)
    p.add_argument("--beta1", type=float, default=None)
    p.add_argument("--beta2", type=float, default=None)
    p.add_argument("--eps", type=float, default=None)
    p.add_argument("--wd", type=float, default=0.2)
    p.add_argument("--warmup", type=int, default=10000)
    p.add_argument("--skip-scheduler", action="store_true", default=False)
    p.add_argument("--lr-scheduler", type=str, default="cosine")
    p.add_argument("--save-frequency", type=int, default=1)
    p.add_argument("--save-most-recent", action="store_true", default=False)
    p.add_argument("--delete-previous-checkpoint", action="store_true", default=False)
    p.add_argument("--zeroshot-frequency", type=int, default=2)
    p.add_argument("--resume", default=None, type=str)
    p.add_argument("--precision", choices=["amp", "amp_bf16", "amp_bfloat16", "bf16", "fp16", "fp32"], default="amp")
    p.add_argument("--model", type=str, default="EVA02-CLIP-B-16")
    p.add_argument("--pretrained", default="", type=str)
    p.add_argument("--lock-image", default=False, action="store_true")
    p.add_argument("--lock-image-unlocked-groups", type=int, default=0)
    p.add_argument("--grad-checkpointing", default=False, action="store_true")
    p.add_argument("--accum-freq", type=int, default=1)
    p.add_argument("--dist-url", default="env://", type=str)
    p.add_argument("--dist-backend", default="nccl", type=str)
    p.add_argument("--log-every-n-steps", type=int, default=100)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--grad-clip-norm", type=float, default=None)
    p.add_argument("--cache-dir", type=str, default="checkpoints")
    p.add_argument("--det-image-size", type=int, default=1024)
    p.add_argument("--downsample-factor", type=int, default=16)
    p.add_argument("--extract-type", default="v2")
    p.add_argument("--alpha", type=float, default=1.0)
    p.add_argument("--cosine-weight", type=float, default=1.0)
    p.add_argument("--multiscale", action="store_true")
    p.add_argument("--crop-scale", type=float, default=1.0)
    p.add_argument("--train-steps-per-epoch", type=int, default=0,
                   help="synthetic_distill only: optimizer steps per epoch (0 = dataset length / global batch)")
    p.add_argument("--synthetic-eval-classes", type=int, default=0,
                   help="synthetic_distill only: > 0 runs the region-classification eval loop (zero_shot.py) on a "
                        "synthetic panoptic-style set with this many classes every --zeroshot-frequency epochs")
    p.add_argument("--image-ave-pool", action="store_true", default=False)
    p.add_argument("--fused-optimizer", action="store_true", default=True)
    args = p.parse_args(args)
    for name, val in get_default_params(args.model).items():
        if getattr(args, name) is None:
            setattr(args, name, val)
    return args
