"""Fused AdamW over the student's flat buffers.

Reproduces the optimizer the reference builds at src/training/main.py:199-213: two parameter groups
(weight decay on >=2-D tensors, none on ndim<2 / bias / ln / logit_scale), torch.optim.AdamW update
rule, parameters whose gradient is None (the last block's q_proj / k_proj / q_bias) are skipped
entirely — they live in the grad-less tail of the flat buffer and are never touched.
"""
from __future__ import annotations

import torch

from . import ops
from .student import StudentEngine


class FusedAdamW:
    def __init__(self, engine: StudentEngine, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.2):
        self.engine = engine
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        n = engine.layout.n_grad
        self.exp_avg = torch.zeros(n, device=engine.device, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=engine.device, dtype=torch.float32)
        self.step_count = 0
        # mirrors of torch's param_groups so the reference's scheduler (`param_group["lr"] = ...`,
        # scheduler.py:4-6) can drive it
        self.param_groups = [dict(lr=lr, weight_decay=0.0, name="no_decay"),
                             dict(lr=lr, weight_decay=weight_decay, name="decay")]

    def zero_grad(self, set_to_none: bool = True) -> None:
        return None          # the backward overwrites the flat gradient buffer every step

    def step(self, grad_scale: float = 1.0) -> None:
        e = self.engine
        if e.grads_ready is not None and e.comm_stream is not None:
            # gradients are being all-reduced on the side stream: update there, right behind the collective
            with torch.cuda.stream(e.comm_stream):
                self._step(grad_scale)
                e.weights_ready = e.comm_stream.record_event()
            return
        self._step(grad_scale)

    def _step(self, grad_scale: float) -> None:
        e, lay = self.engine, self.engine.layout
        self.step_count += 1
        b1, b2 = self.betas
        nd, ng = lay.n_decay, lay.n_grad
        ds, ns = lay.decay_start(e.first_trainable), lay.nodecay_start(e.first_trainable)   # frozen blocks are skipped
        g_decay, g_nodecay = self.param_groups[1], self.param_groups[0]
        ops.adamw_step(e.flat_param[ds:nd], e.flat_grad[ds:nd], self.exp_avg[ds:nd], self.exp_avg_sq[ds:nd],
                       g_decay["lr"], b1, b2, self.eps, g_decay["weight_decay"], self.step_count, grad_scale)
        ops.adamw_step(e.flat_param[ns:ng], e.flat_grad[ns:ng], self.exp_avg[ns:ng], self.exp_avg_sq[ns:ng],
                       g_nodecay["lr"], b1, b2, self.eps, g_nodecay["weight_decay"], self.step_count, grad_scale)
        e.weights_epoch += 1         # invalidates forward-only packs built from these weights (model.py: _infer_engine_native)

    def state_dict(self):
        return dict(exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, step=self.step_count,
                    param_groups=self.param_groups)

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_count = int(sd["step"])
        self.param_groups = sd["param_groups"]
