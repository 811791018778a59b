"""CLIPSelf distillation step — drop-in for `training.clipself.CLIPSelf` (src/training/clipself.py:6-49).

Same call signature and return value as the reference plug-in invoked at train.py:91-92:

    losses, batch_size, logit_scale = method(batch, model, dist_model, loss, device, cast_dtype,
                                             distributed, args)

Differences in HOW (not what):
  * index extraction (clipself.py:29-36) is done once on the host when the batch arrives as CPU
    tensors (no per-image device syncs, only valid crops cross PCIe) or by one CUDA kernel when the
    batch is already device resident;
  * teacher forward, student dense forward, RoIAlign, normalise + cosine loss and the whole
    backward run in the sm_100a CUDA library;
  * with `distributed=True` the student gradient is mean-all-reduced once per step over the flat
    gradient buffer (the reference's DDP wrapper is bypassed by `model.module`).
"""
from __future__ import annotations

import random

import torch

from .. import ops
from ..crops import RawImageBatch, _BatchCropper
from ..model import cosine_distill_loss


class CLIPSelf:
    def __init__(self):
        self._copy_stream = None
        self._crops_dev = None
        self._crops_free = None
        self._images_dev = None
        self._images_free = None
        self._cropper = None
        self._teacher_out = None

    def _stream_inputs(self, images, image_crops, valid, R, device, dtype, pieces):
        """Host->device copies of one step on a side stream, in the order the step consumes them: the
        student images first, then only the valid crops, one event per teacher piece, so the PCIe transfer
        overlaps the student forward and the earlier teacher pieces.  Runs of consecutive valid rows go as
        single cudaMemcpyAsync calls straight from the (pinned) batch tensor: no host-side gather."""
        B, K = valid.shape
        flat = image_crops.reshape(B * K, *image_crops.shape[2:])
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=device)
            self._crops_free = torch.cuda.Event()
            self._images_free = torch.cuda.Event()
        cs = self._copy_stream
        if self._images_dev is None or self._images_dev.shape != images.shape or self._images_dev.dtype != dtype:
            self._images_dev = torch.empty(images.shape, device=device, dtype=dtype)
            # a fresh block from the caching allocator may still be read by kernels queued on the compute
            # stream (the allocator only orders reuse within that stream): order the copies after them
            cs.wait_stream(torch.cuda.current_stream())
        else:
            cs.wait_event(self._images_free)            # previous step's student forward has read them
        if self._crops_dev is None or self._crops_dev.shape[0] < R or self._crops_dev.shape[1:] != flat.shape[1:] \
                or self._crops_dev.dtype != dtype:
            self._crops_dev = torch.empty((B * K,) + tuple(flat.shape[1:]), device=device, dtype=dtype)
            cs.wait_stream(torch.cuda.current_stream())
        else:
            cs.wait_event(self._crops_free)             # previous step's teacher is done reading
        dev = self._crops_dev
        idx = valid.flatten().nonzero().flatten().tolist()
        events = []
        with torch.cuda.stream(cs):
            self._images_dev.copy_(images, non_blocking=True)
            images_ready = torch.cuda.Event()
            images_ready.record(cs)
            for s0, n in pieces:
                a = s0
                while a < s0 + n:                               # maximal run of consecutive source rows
                    b = a + 1
                    while b < s0 + n and idx[b] == idx[b - 1] + 1:
                        b += 1
                    dev[a:b].copy_(flat[idx[a]:idx[a] + (b - a)], non_blocking=True)
                    a = b
                ev = torch.cuda.Event()
                ev.record(cs)
                events.append(ev)
        return self._images_dev, images_ready, dev[:R], events

    def __call__(self, batch, model, dist_model, loss, device, cast_dtype, distributed, args):
        if distributed:
            model = getattr(model, "module", model)
            dist_model = getattr(dist_model, "module", dist_model)
        device = torch.device(device)
        dtype = cast_dtype if cast_dtype is not None else torch.float32
        crop_events = None
        streamed_images = False
        raw = batch if isinstance(batch, RawImageBatch) else None
        if raw is not None:
            # image-backed dataset: decoded uint8 images + boxes arrive, the student images (ResizeLongest + pad) and the K
            # teacher crops per image (crop -> ResizeMaxSize bicubic -> pad -> normalise; data.py:226-245, transform.py:26-49,
            # 169-191) are produced on the device, bit-exact with the reference's PIL path.  Index extraction as below.
            normed_boxes = raw.normed_boxes
            B, K = normed_boxes.shape[:2]
            boxes32 = normed_boxes.float()
            valid = boxes32[:, :, 4] > 0.5
            offsets = torch.zeros(B + 1, dtype=torch.int32)
            offsets[1:] = valid.sum(1).cumsum(0)
            R = int(offsets[-1])
            assert [len(b) for b in raw.crop_boxes_px] == valid.sum(1).tolist(), "one crop rectangle per valid box"
            rois = boxes32[valid][:, :4].contiguous().to(device, non_blocking=True)
            offsets = offsets.to(device, non_blocking=True)
            if self._cropper is None:
                self._cropper = _BatchCropper()
            images, crops = self._cropper(raw, device)
            crops = self._cropper.cast(crops, dtype)
        else:
            images, normed_boxes, image_crops = batch       # texts are not paired with images
            B, K = normed_boxes.shape[:2]
        if raw is not None:
            pass                                        # images / rois / crops are on the device already
        elif normed_boxes.device.type == "cpu":
            # host-side, bit-exact: valid = boxes[..., 4] > 0.5, image-major order (clipself.py:29-36)
            boxes32 = normed_boxes.float()
            valid = boxes32[:, :, 4] > 0.5
            counts = valid.sum(1)
            offsets = torch.zeros(B + 1, dtype=torch.int32)
            offsets[1:] = counts.cumsum(0)
            R = int(offsets[-1])
            rois = boxes32[valid][:, :4].contiguous().to(device, non_blocking=True)
            offsets = offsets.to(device, non_blocking=True)
            if images.device.type == "cpu":
                images, images_ready, crops, crop_events = self._stream_inputs(
                    images, image_crops, valid, R, device, dtype, dist_model.visual.teacher_chunk_schedule(R))
                torch.cuda.current_stream().wait_event(images_ready)
                streamed_images = True
            else:
                raise ValueError("boxes on the host but images on the device: pass the whole batch on one side")
        else:
            rois_all, crop_index, _, offsets = ops.extract_rois(normed_boxes.float().contiguous())
            R = int(offsets[-1])                         # the step's single device->host sync
            rois = rois_all[:R]
            flat = image_crops.reshape(B * K, *image_crops.shape[2:])
            crops = ops.gather_rows(flat.contiguous(), crop_index, R) if R != B * K else flat
            crops = crops.to(dtype)
        if not streamed_images:
            images = images.to(device=device, dtype=dtype, non_blocking=True)

        if getattr(args, "multiscale", False):
            cur_h, cur_w = images.shape[2:]
            assert cur_h == cur_w
            if cur_h == 1024:
                tar_sizes = [320, 640, 896, 1024]
            elif cur_h == 896:
                tar_sizes = [336, 448, 672, 896]
            else:
                raise NotImplementedError
            tar_size = random.choice(tar_sizes)
            if tar_size != cur_h:
                images = ops.resize_bilinear(images.contiguous(), tar_size)

        model.visual.sync_gradients = bool(distributed)

        def run_teacher():
            with torch.no_grad():
                # the tower writes into a buffer this plug-in owns (a stable address keeps the native tower's per-buffer CUDA
                # graphs hot); the loss gets a private copy (4 MB), so a second call before backward() cannot change what
                # autograd saved
                edim = dist_model.visual.cfg.embed_dim
                if self._teacher_out is None or self._teacher_out.shape[0] < R or self._teacher_out.device != device:
                    self._teacher_out = torch.empty(max(R, 1), edim, device=device, dtype=torch.float32)
                feats = dist_model.visual.forward_chunked(crops, crop_events, out=self._teacher_out[:R])
                if crop_events is not None:
                    self._crops_free.record()
                return feats.clone()

        # Order: normally the student forward goes first (it only needs the small images, so it overlaps the crop H2D
        # stream).  With the overlapped gradient exchange the frozen teacher goes first: it does not depend on the weights
        # the previous step's all-reduce + AdamW are still producing on the side stream, the student forward does.
        teacher_first = bool(distributed) and model.visual.overlap_gradient_sync
        teacher_crop_features = run_teacher() if teacher_first else None
        student_roi_features = model.visual.roi_features_packed(images, rois, offsets, R)
        if streamed_images:
            self._images_free.record()
        if teacher_crop_features is None:
            teacher_crop_features = run_teacher()

        loss_cosine = cosine_distill_loss(student_roi_features, teacher_crop_features,
                                          float(getattr(args, "cosine_weight", 1.0)))
        losses = dict(loss_cosine=loss_cosine)
        return losses, len(images), model.logit_scale.exp()
