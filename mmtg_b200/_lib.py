"""ctypes binding of libmmtg_b200.so (the C-ABI declared in include/mmtg_b200.h).

PyTorch is only used by callers for device memory and streams: every call here passes raw
device pointers (`tensor.data_ptr()`) plus the current CUDA stream handle. There is no CPU
fallback — if the library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmmtg_b200.so")

ACT_NONE, ACT_TANH, ACT_GELU_NEW = 0, 1, 2
F32, BF16 = 0, 1


class MMTGError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64),
        ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("split_k", C.c_int32), ("block_n", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("out_dtype", C.c_int32), ("accumulate", C.c_int32),
        ("out2", C.c_void_p), ("ldo2", C.c_int64),
        ("bias", C.c_void_p), ("act", C.c_int32), ("_pad0", C.c_int32),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("dgelu_src", C.c_void_p), ("ldg", C.c_int64),
        ("rowtab0", C.c_void_p), ("rowidx0", C.c_void_p), ("ldt0", C.c_int64),
        ("rowmod0", C.c_int32), ("_pad1", C.c_int32),
        ("rowtab1", C.c_void_p), ("rowidx1", C.c_void_p), ("ldt1", C.c_int64),
        ("colsum", C.c_void_p),
        ("lse_partial", C.c_void_p),
        ("dact_tanh_out", C.c_int32), ("out2_mode", C.c_int32),
        ("drop_seed", C.c_void_p), ("drop_site", C.c_uint32), ("drop_p", C.c_float),
        ("grid_mode", C.c_int32), ("_pad2", C.c_int32),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load the shared library once; fail loudly when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MMTGError(
                f"{LIB_PATH} not found — build it with `make` (or __graft_entry__.build()); "
                "mmtg_b200 has no CPU or PyTorch fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.mmtg_last_error.restype = C.c_char_p
        _lib.mmtg_launch_count.restype = C.c_int64
        _lib.mmtg_train_workspace_bytes.restype = C.c_int64
        _lib.mmtg_decode_workspace_bytes.restype = C.c_int64
        _lib.mmtg_ws_lse.restype = C.c_void_p
        _lib.mmtg_ws_dlogits_bf16.restype = C.c_void_p
        if _lib.mmtg_abi_version() != 3:
            raise MMTGError("libmmtg_b200.so ABI version mismatch")
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mmtg_last_error().decode("utf-8", "replace")
        raise MMTGError(f"{what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(lib().mmtg_launch_count())


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()
