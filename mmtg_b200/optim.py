"""Fused optimizer step over the model's flat buffers.

`FusedAdamW` restates what the reference trains with — `clip_grad_norm_(params, 1.0)` followed by
`transformers.AdamW(lr)` (src/train.py:137,194-195: betas .9/.999, eps 1e-6 outside the
bias-corrected denominator, weight decay 0) — as two kernels over ONE contiguous buffer, and
refreshes the bf16 weight shadow in the same pass. It subclasses torch.optim.Optimizer only so
LR schedulers (`get_linear_schedule_with_warmup`, src/train.py:146) can drive `param_groups`.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-5, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0,
                 correct_bias=True, max_grad_norm=None):
        self.model = model
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias)
        super().__init__(list(model.parameters()), defaults)
        self.max_grad_norm = max_grad_norm
        self._t = 0
        self._m = self._v = None
        self._partial = self._normsq = None

    def _buffers(self):
        P, W16, G = self.model._flat
        if self._m is None or self._m.device != P.device or self._m.numel() != P.numel():
            self._m, self._v = torch.zeros_like(P), torch.zeros_like(P)
            self._partial = torch.empty(1024, device=P.device)
            self._normsq = torch.zeros(1, device=P.device)
            self._lr_dev = torch.zeros(1, device=P.device)
            self._step_dev = torch.zeros(1, device=P.device, dtype=torch.int32)
            self._lr_host = None
        return P, W16, G

    @torch.no_grad()
    def step(self, closure=None):
        if self.model._flat is None:
            raise _lib.MMTGError("FusedAdamW.step() before the first forward/backward")
        P, W16, G = self._buffers()
        lib, st = _lib.lib(), C.c_void_p(_lib.stream_ptr())
        g = self.param_groups[0]
        self._t += 1
        if self._lr_host != g["lr"] and not torch.cuda.is_current_stream_capturing():
            self._lr_dev.fill_(g["lr"])  # schedule state lives on the device (graph-replayable)
            self._lr_host = g["lr"]
        normsq = None
        if self.max_grad_norm is not None:
            _lib.check(lib.mmtg_grad_norm_sq(C.c_void_p(G.data_ptr()), C.c_int64(G.numel()),
                                             C.c_void_p(self._partial.data_ptr()), 1024,
                                             C.c_void_p(self._normsq.data_ptr()), st), "mmtg_grad_norm_sq")
            normsq = C.c_void_p(self._normsq.data_ptr())
        _lib.check(lib.mmtg_adamw_step(C.c_void_p(P.data_ptr()), C.c_void_p(G.data_ptr()),
                                       C.c_void_p(self._m.data_ptr()), C.c_void_p(self._v.data_ptr()),
                                       C.c_void_p(W16.data_ptr()), C.c_int64(P.numel()), C.c_float(g["lr"]),
                                       C.c_double(g["betas"][0]), C.c_double(g["betas"][1]), C.c_float(g["eps"]),
                                       C.c_float(g["weight_decay"]), self._t, int(g["correct_bias"]), normsq,
                                       C.c_float(self.max_grad_norm or 0.0), C.c_void_p(self._lr_dev.data_ptr()),
                                       C.c_void_p(self._step_dev.data_ptr()), st), "mmtg_adamw_step")
        self.model.mark_bf16_shadow_fresh()  # the kernel rewrote the bf16 weight shadow itself
        return None

    def sync_lr(self):
        """Push `param_groups[0]['lr']` (as an LR scheduler leaves it) to the device copy the
        captured AdamW kernel reads. GraphedTrainStep calls this before every replay."""
        lr = self.param_groups[0]["lr"]
        if self._lr_host != lr and self._m is not None:
            self._lr_dev.fill_(lr)
            self._lr_host = lr

    def state_dict(self):
        """torch.optim.Optimizer.state_dict() plus the flat moments and the device step counter
        (they live outside `self.state`), so save/resume keeps exp_avg, exp_avg_sq and the bias
        correction."""
        sd = super().state_dict()
        if self._m is not None:
            sd["mmtg_flat"] = {"exp_avg": self._m.detach().cpu(), "exp_avg_sq": self._v.detach().cpu(),
                               "step": int(self._step_dev.item())}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flat = state_dict.pop("mmtg_flat", None)
        super().load_state_dict(state_dict)
        if flat is not None:
            if self.model._flat is None:
                raise _lib.MMTGError("FusedAdamW.load_state_dict: move the model to its CUDA device and run one "
                                     "forward (or call model._ensure_flat(device)) first")
            self._buffers()
            self._m.copy_(flat["exp_avg"])
            self._v.copy_(flat["exp_avg_sq"])
            self._step_dev.fill_(int(flat["step"]))
            self._t = int(flat["step"])
            self._lr_host = None

    def set_lr(self, lr):
        """Update the learning rate (host + device copy); safe between CUDA-graph replays."""
        for g in self.param_groups:
            g["lr"] = lr
        self._buffers()
        self._lr_dev.fill_(lr)
        self._lr_host = lr

    def grad_norm(self):
        """Global L2 norm computed by the last step() (device tensor)."""
        return self._normsq.sqrt()

    def zero_grad(self, set_to_none: bool = False):
        if self.model._flat is None or set_to_none:
            return super().zero_grad(set_to_none=True)
        self.model._flat[2].zero_()
