"""Times the in-step epilogue variants of the tcgen05 GEMM on the decoder shapes (CUDA events,
L2 flushed between launches): the kernels the train step actually runs, not the bare product."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mmtg_b200 import _lib, ops

dev = torch.device("cuda:0")
M, E = 7552, 768
bf = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def rnd(*s, dt=bf):
    return (torch.randn(*s, device=dev) * 0.5).to(dt)


def time_it(fn, n=7):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


x, att, a4, g16, du = rnd(M, E), rnd(M, E), rnd(M, 4 * E), rnd(M, E), rnd(M, 4 * E)
w_attn, w_proj, w_fc, w_proj2 = rnd(E, 3 * E), rnd(E, E), rnd(E, 4 * E), rnd(4 * E, E)
b3, b1, b4 = rnd(3 * E, dt=torch.float32), rnd(E, dt=torch.float32), rnd(4 * E, dt=torch.float32)
h = rnd(M, E, dt=torch.float32)
qkv, a_out, u_out = torch.empty(M, 3 * E, device=dev, dtype=bf), torch.empty(M, 4 * E, device=dev, dtype=bf), torch.empty(M, 4 * E, device=dev, dtype=bf)
h_out = torch.empty(M, E, device=dev)
dx = torch.empty(M, E, device=dev, dtype=bf)
du_out = torch.empty(M, 4 * E, device=dev, dtype=bf)
cs = torch.zeros(4 * E, device=dev)
seed = torch.tensor([7], dtype=torch.int64, device=dev)

cases = {
    "c_attn fwd (bias)": (lambda: ops.gemm(x, w_attn, qkv, M=M, N=3 * E, K=E, b_mn_major=True, bias=b3), 3 * E * E),
    "c_fc fwd (bias+gelu+deriv out2)": (lambda: ops.gemm(x, w_fc, a_out, M=M, N=4 * E, K=E, b_mn_major=True, bias=b4, act=ops.ACT_GELU_NEW, out2=u_out, out2_mode=1), 4 * E * E),
    "attn c_proj fwd (bias+res)": (lambda: ops.gemm(att, w_proj, h_out, M=M, N=E, K=E, b_mn_major=True, bias=b1, residual=h), E * E),
    "attn c_proj fwd (bias+res+drop)": (lambda: ops.gemm(att, w_proj, h_out, M=M, N=E, K=E, b_mn_major=True, bias=b1, residual=h, drop=(seed, 5, 0.1)), E * E),
    "mlp c_proj fwd (bias+res)": (lambda: ops.gemm(a4, w_proj2, h_out, M=M, N=E, K=4 * E, b_mn_major=True, bias=b1, residual=h), 4 * E * E),
    "mlp c_proj fwd (bias+res+drop)": (lambda: ops.gemm(a4, w_proj2, h_out, M=M, N=E, K=4 * E, b_mn_major=True, bias=b1, residual=h, drop=(seed, 6, 0.1)), 4 * E * E),
    "du dgrad (dmul+colsum)": (lambda: ops.gemm(g16, w_proj2, du_out, M=M, N=4 * E, K=E, dgelu_src=u_out, dact_mode=2, colsum=cs), 4 * E * E),
    "du dgrad (dmul only)": (lambda: ops.gemm(g16, w_proj2, du_out, M=M, N=4 * E, K=E, dgelu_src=u_out, dact_mode=2), 4 * E * E),
    "du dgrad (colsum only)": (lambda: ops.gemm(g16, w_proj2, du_out, M=M, N=4 * E, K=E, colsum=cs), 4 * E * E),
    "du dgrad shape, plain": (lambda: ops.gemm(g16, w_proj2, du_out, M=M, N=4 * E, K=E), 4 * E * E),
    "c_fc dgrad (plain)": (lambda: ops.gemm(du, w_fc, dx, M=M, N=E, K=4 * E), 4 * E * E),
    "c_attn dgrad (plain)": (lambda: ops.gemm(qkv, w_attn, dx, M=M, N=E, K=3 * E), 3 * E * E),
}
for bn in (128, 256):
    cases[f"attn c_proj fwd (bias+res+drop) bn={bn}"] = (lambda bn=bn: ops.gemm(att, w_proj, h_out, M=M, N=E, K=E, b_mn_major=True, bias=b1, residual=h, drop=(seed, 5, 0.1), block_n=bn), E * E)
    cases[f"mlp c_proj fwd (bias+res+drop) bn={bn}"] = (lambda bn=bn: ops.gemm(a4, w_proj2, h_out, M=M, N=E, K=4 * E, b_mn_major=True, bias=b1, residual=h, drop=(seed, 6, 0.1), block_n=bn), 4 * E * E)
    cases[f"c_fc dgrad (plain) bn={bn}"] = (lambda bn=bn: ops.gemm(du, w_fc, dx, M=M, N=E, K=4 * E, block_n=bn), 4 * E * E)
    cases[f"c_attn dgrad (plain) bn={bn}"] = (lambda bn=bn: ops.gemm(qkv, w_attn, dx, M=M, N=E, K=3 * E, block_n=bn), 3 * E * E)
    cases[f"attn proj dgrad (plain) bn={bn}"] = (lambda bn=bn: ops.gemm(g16, w_proj, dx, M=M, N=E, K=E, block_n=bn), E * E)
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, (fn, nk) in cases.items():
    if only and only not in name:
        continue
    ms = time_it(fn)
    print(f"{name:36s} {ms * 1e3:7.1f} us  {2 * M * nk / ms / 1e9:6.0f} TFLOP/s", flush=True)
