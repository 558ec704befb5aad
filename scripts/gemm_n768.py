"""The decoder's N = 768 GEMM shapes at the bench batch: tile width 256 vs 192 vs cuBLAS."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmtg_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=7):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3
M, N = 7552, 768
for K in (768, 2304, 3072):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    r = {bn: timeit(lambda: ops.gemm(A, B, out, M=M, N=N, K=K, block_n=bn)) for bn in (128, 192, 256)}
    cb = timeit(lambda: torch.matmul(A, B.t(), out=out))
    print(f"{M}x{N}x{K}: " + "  ".join(f"bn{bn} {t:6.1f} us" for bn, t in r.items()) + f"  cuBLAS {cb:6.1f} us", flush=True)
