"""K sweep of the tcgen05 GEMM at the dominant M x N (7552 x 3072) next to cuBLAS (torch.matmul):
separates the steady-state mainloop rate from the per-tile epilogue pace."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmtg_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=5):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


shapes = [(7552, 3072, k) for k in (256, 768, 1536, 3072, 6144)] + [(7552, 2304, 768), (7552, 768, 3072), (8192, 8192, 8192),
                                                                      (7552, 13320, 768)]
for M, N, K in shapes:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.gemm(A, B, out, M=M, N=N, K=K, block_n=256))
    Bt = B.t()
    ms_cb = timeit(lambda: torch.matmul(A, Bt, out=out))
    print(f"{M}x{N}x{K}: ours {ms*1e3:8.1f} us {2*M*N*K/ms/1e9:6.0f} TF | cuBLAS {ms_cb*1e3:8.1f} us {2*M*N*K/ms_cb/1e9:6.0f} TF", flush=True)
