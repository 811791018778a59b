#!/bin/bash
# the reference's launch line (scripts/train_clipself_coco_image_patches_eva_vitb16.sh) on synthetic data:
# 4 optimizer steps with the student at 448 px (785 tokens), gradient clipping, the epoch-end student/teacher
# ensemble checkpoint and one pass of the region-classification eval loop (added late in round 1: the eval /
# clipping lines have not run on a GPU yet — run this first in round 2)
cd "$(dirname "$0")/.."
PYTHONPATH=.:compat python -m training.main --batch-size 8 --lr 1e-5 --wd 0.1 --epochs 1 --workers 0 \
  --model EVA02-CLIP-B-16 --pretrained eva --warmup 2 --zeroshot-frequency 1 --dataset-type synthetic_distill \
  --cache-dir "" --log-every-n-steps 1 --lock-image --save-frequency 1 --lock-image-unlocked-groups 12 \
  --extract-type="v2" --name smoke --downsample-factor 16 --det-image-size 448 --alpha 0.7 --max-boxes 8 \
  --train-steps-per-epoch 4 --logs /tmp/clipself_smoke_logs --grad-clip-norm 5.0 --synthetic-eval-classes 7
