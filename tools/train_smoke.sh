#!/bin/bash
# the reference's launch line (scripts/train_clipself_coco_image_patches_eva_vitb16.sh) on synthetic data
cd "$(dirname "$0")/.."
PYTHONPATH=.:compat python -m training.main --batch-size 8 --lr 1e-5 --wd 0.1 --epochs 1 --workers 0 \
  --model EVA02-CLIP-B-16 --pretrained eva --warmup 2 --zeroshot-frequency 1 --dataset-type synthetic_distill \
  --cache-dir "" --log-every-n-steps 1 --lock-image --save-frequency 1 --lock-image-unlocked-groups 12 \
  --extract-type="v2" --name smoke --downsample-factor 16 --det-image-size 224 --alpha 0.7 --max-boxes 8 \
  --train-steps-per-epoch 4 --logs /tmp/clipself_smoke_logs
