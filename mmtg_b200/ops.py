"""Thin Python wrappers over the C-ABI ops (raw pointers + current stream)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ACT_GELU_NEW, ACT_NONE, ACT_TANH, BF16, F32  # noqa: F401


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "matrix must be row-major with unit inner stride"
    return t.stride(0)


def gemm(A: torch.Tensor, B: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int,
         a_mn_major: bool = False, b_mn_major: bool = False, bias=None, act: int = ACT_NONE,
         out2=None, residual=None, dgelu_src=None, rowtab0=None, rowidx0=None, rowmod0: int = 0,
         rowtab1=None, rowidx1=None, colsum=None, lse_partial=None, accumulate: bool = False,
         split_k: int = 1, block_n: int = 0, skinny: bool = False, out2_mode: int = 0, dact_mode: int = 0,
         drop=None) -> torch.Tensor:
    """out = epilogue(A · Bᵀ) on the tcgen05 GEMM (see include/mmtg_b200.h: mmtg_gemm_bf16)."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    assert out.dtype in (torch.float32, torch.bfloat16)
    a = _lib.GemmArgs()
    a.A, a.B = A.data_ptr(), B.data_ptr()
    a.lda, a.ldb = _ld(A), _ld(B)
    a.a_mn_major, a.b_mn_major = int(a_mn_major), int(b_mn_major)
    a.M, a.N, a.K = M, N, K
    a.split_k, a.block_n = split_k, block_n
    a.out, a.ldo = out.data_ptr(), _ld(out)
    a.out_dtype = F32 if out.dtype == torch.float32 else BF16
    a.accumulate = int(accumulate)
    if out2 is not None:
        assert out2.dtype == torch.bfloat16
        a.out2, a.ldo2 = out2.data_ptr(), _ld(out2)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= N
        a.bias = bias.data_ptr()
    a.act = act
    if residual is not None:
        assert residual.dtype == torch.float32
        a.residual, a.ldr = residual.data_ptr(), _ld(residual)
    if dgelu_src is not None:
        assert dgelu_src.dtype == torch.bfloat16
        a.dgelu_src, a.ldg = dgelu_src.data_ptr(), _ld(dgelu_src)
    if rowtab0 is not None:
        assert rowtab0.dtype == torch.float32
        a.rowtab0, a.ldt0 = rowtab0.data_ptr(), _ld(rowtab0)
        if rowidx0 is not None:
            assert rowidx0.dtype == torch.int32
            a.rowidx0 = rowidx0.data_ptr()
        a.rowmod0 = rowmod0
    if rowtab1 is not None:
        assert rowtab1.dtype == torch.float32 and rowidx1.dtype == torch.int32
        a.rowtab1, a.ldt1, a.rowidx1 = rowtab1.data_ptr(), _ld(rowtab1), rowidx1.data_ptr()
    if colsum is not None:
        assert colsum.dtype == torch.float32
        a.colsum = colsum.data_ptr()
    if lse_partial is not None:
        assert lse_partial.dtype == torch.float32
        a.lse_partial = lse_partial.data_ptr()
    a.out2_mode, a.dact_tanh_out = out2_mode, dact_mode
    if drop is not None:  # (device int64 seed tensor, site, p)
        a.drop_seed, a.drop_site, a.drop_p = drop[0].data_ptr(), int(drop[1]), float(drop[2])
    fn = _lib.lib().mmtg_skinny_gemm_bf16 if skinny else _lib.lib().mmtg_gemm_bf16
    _lib.check(fn(C.byref(a), C.c_void_p(_lib.stream_ptr())), "mmtg_gemm_bf16")
    return out


def _st():
    return C.c_void_p(_lib.stream_ptr())


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def layernorm_fwd(x, gamma, beta, eps=1e-5, want_bf16=True, want_f32=False):
    M, E = x.shape
    y16 = torch.empty(M, E, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    y32 = torch.empty(M, E, device=x.device) if want_f32 else None
    mean, rstd = torch.empty(M, device=x.device), torch.empty(M, device=x.device)
    _lib.check(_lib.lib().mmtg_layernorm_fwd(_p(x), _p(gamma), _p(beta), _p(y16), _p(y32), _p(mean), _p(rstd),
                                             M, E, C.c_float(eps), _st()), "mmtg_layernorm_fwd")
    return y16, y32, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dx, accumulate, dgamma, dbeta, dx16=None, dx_colsum=None):
    M, E = x.shape
    _lib.check(_lib.lib().mmtg_layernorm_bwd(_p(dy), int(dy.dtype == torch.bfloat16), _p(x), _p(mean), _p(rstd),
                                             _p(gamma), _p(dx), int(accumulate), _p(dgamma), _p(dbeta),
                                             _p(dx16), _p(dx_colsum), M, E, _st()), "mmtg_layernorm_bwd")


def ln_param_grads(dy16, x, mean, rstd, dx16, dgamma, dbeta, dx_colsum):
    M, E = x.shape
    _lib.check(_lib.lib().mmtg_ln_param_grads(_p(dy16), _p(x), _p(mean), _p(rstd), _p(dx16), _p(dgamma), _p(dbeta),
                                              _p(dx_colsum), M, E, _st()), "mmtg_ln_param_grads")


def colsum(x, out, copy16=None):
    M, N = x.shape
    _lib.check(_lib.lib().mmtg_colsum(_p(x), int(x.dtype == torch.bfloat16), C.c_int64(x.stride(0)), _p(copy16),
                                      C.c_int64(copy16.stride(0) if copy16 is not None else 0), _p(out), M, N,
                                      _st()), "mmtg_colsum")


def attn_fwd(qkv, mask, B, L, NH, impl=0):
    """impl: 0 default, 1 mma.sync tiles, 2 tcgen05/TMEM."""
    E = NH * 64
    out = torch.empty(B * L, E, device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty(B, NH, L, device=qkv.device)
    _lib.check(_lib.lib().mmtg_attn_fwd_ex(_p(qkv), _p(mask), _p(out), _p(lse), B, L, NH, impl, _st()),
               "mmtg_attn_fwd")
    return out, lse


def attn_bwd(qkv, mask, out, dout, lse, B, L, NH, impl=0):
    dqkv = torch.full_like(qkv, float("nan"))
    delta = torch.empty(B, NH, L, device=qkv.device)
    _lib.check(_lib.lib().mmtg_attn_bwd_ex(_p(qkv), _p(mask), _p(out), _p(dout), _p(lse), _p(delta), _p(dqkv), B,
                                           L, NH, impl, _st()), "mmtg_attn_bwd")
    return dqkv
