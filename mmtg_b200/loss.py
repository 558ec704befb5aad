"""Drop-in `MyLoss` (mirrors /root/reference/src/loss.py:39-74): curriculum negative-sampling loss.

Two device paths, both hand-written kernels (mmtg_b200/csrc/loss.cu), no PyTorch fallback:
  * fused  — `outputs` is the logits tensor returned by `mmtg_b200.MMTG.forward`: the row
    log-sum-exp already produced by the lm_head GEMM epilogue is reused, and backward writes the
    bf16 dlogits operand of the lm_head dgrad/wgrad GEMMs directly (no fp32 [B,L,V] gradient);
  * generic — any fp32 logits tensor: one pass for the row LSE, dense fp32 gradient on backward.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _vp(p):
    return C.c_void_p(p)


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, token, step, targets, ratings, stage):
        lib = _lib.lib()
        d = step.dims
        dev = logits.device
        st = _vp(_lib.stream_ptr())
        ce = torch.empty(d.B, device=dev)
        coef = torch.empty(d.B, device=dev)
        loss = torch.empty((), device=dev)
        _lib.check(lib.mmtg_ce_reduce(_vp(logits.data_ptr()), C.c_int64(d.V), _vp(step.lse_ptr), None,
                                      _vp(targets.data_ptr()), None, _vp(ce.data_ptr()), None, d.B, d.L,
                                      d.P, d.T, st), "mmtg_ce_reduce")
        _lib.check(lib.mmtg_negloss(_vp(ce.data_ptr()), _vp(ratings.data_ptr()), stage, _vp(loss.data_ptr()),
                                    _vp(coef.data_ptr()), d.B, st), "mmtg_negloss")
        ctx.step = step  # holds the logits only weakly (model._Step): no cycle through this node
        ctx.save_for_backward(logits, coef, targets)
        return loss

    @staticmethod
    def backward(ctx, g):
        step, d = ctx.step, ctx.step.dims
        logits, coef, targets = ctx.saved_tensors
        if step.serial != step.model._serial:
            raise _lib.MMTGError("loss.backward() after a newer forward(): activation workspace overwritten")
        g = g.detach().float().contiguous()
        _lib.check(_lib.lib().mmtg_ce_bwd(_vp(logits.data_ptr()), C.c_int64(d.V), _vp(step.lse_ptr), None,
                                          _vp(targets.data_ptr()), _vp(coef.data_ptr()),
                                          _vp(g.data_ptr()), None, _vp(step.dlogits_ptr), 1, C.c_int64(d.Vp),
                                          d.B, d.L, d.P, d.T, d.V, _vp(_lib.stream_ptr())), "mmtg_ce_bwd")
        step.dlogits_ready = True
        return None, torch.zeros_like(g).reshape(()), None, None, None, None


class _GenericLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, ratings, stage, P):
        lib = _lib.lib()
        B, L, V = logits.shape
        T = targets.shape[1]
        dev = logits.device
        st = _vp(_lib.stream_ptr())
        lse = torch.empty(B * L, device=dev)
        ce = torch.empty(B, device=dev)
        coef = torch.empty(B, device=dev)
        loss = torch.empty((), device=dev)
        _lib.check(lib.mmtg_lse_rows(_vp(logits.data_ptr()), C.c_int64(V), _vp(lse.data_ptr()), B * L, V, st),
                   "mmtg_lse_rows")
        _lib.check(lib.mmtg_ce_reduce(_vp(logits.data_ptr()), C.c_int64(V), _vp(lse.data_ptr()), None,
                                      _vp(targets.data_ptr()), None, _vp(ce.data_ptr()), None, B, L, P, T, st),
                   "mmtg_ce_reduce")
        _lib.check(lib.mmtg_negloss(_vp(ce.data_ptr()), _vp(ratings.data_ptr()), stage, _vp(loss.data_ptr()),
                                    _vp(coef.data_ptr()), B, st), "mmtg_negloss")
        ctx.save_for_backward(logits, lse, coef, targets)
        ctx.P = P
        return loss

    @staticmethod
    def backward(ctx, g):
        logits, lse, coef, targets = ctx.saved_tensors
        B, L, V = logits.shape
        T = targets.shape[1]
        g = g.detach().float().contiguous()
        out = torch.empty_like(logits)
        _lib.check(_lib.lib().mmtg_ce_bwd(_vp(logits.data_ptr()), C.c_int64(V), _vp(lse.data_ptr()), None,
                                          _vp(targets.data_ptr()), _vp(coef.data_ptr()), _vp(g.data_ptr()),
                                          None, _vp(out.data_ptr()), 0, C.c_int64(V), B, L, ctx.P, T, V,
                                          _vp(_lib.stream_ptr())), "mmtg_ce_bwd")
        return out, None, None, None, None


class MyLoss(torch.nn.Module):
    def __init__(self, data_config, model_cfgs):
        super().__init__()
        self._max_topic_len = data_config.topic_prompt_length
        self._seq_len = model_cfgs["seq_len"]

    def forward(self, outputs, targets, ratings, stage):
        """outputs [B, P+T, V] fp32 logits; targets [B, T]; ratings [B] in 1..5; stage 1|2|3."""
        if not outputs.is_cuda:
            raise _lib.MMTGError("mmtg_b200.MyLoss runs on CUDA only (no CPU fallback)")
        t32 = targets.to(device=outputs.device, dtype=torch.int32).contiguous()
        r32 = ratings.to(device=outputs.device, dtype=torch.int32).contiguous()
        step = getattr(outputs, "_mmtg_step", None)
        fused = (step is not None and step.serial == step.model._serial and outputs is step.logits
                 and getattr(step, "token", None) is not None and outputs.requires_grad
                 and torch.is_grad_enabled() and t32.shape[1] == step.T)
        if fused:
            return _FusedLoss.apply(outputs, step.token, step, t32, r32, int(stage))
        if outputs.dtype != torch.float32 or not outputs.is_contiguous():
            outputs = outputs.float().contiguous()
        assert outputs.shape[1] == self._max_topic_len + t32.shape[1]
        return _GenericLoss.apply(outputs, t32, r32, int(stage), self._max_topic_len)
