"""First-contact probe for the tcgen05 GEMM: each layout variant in its own process so a trap
in one does not poison the others. Prints max |err| per variant and a quick timing."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, torch, math
sys.path.insert(0, %r)
from mmtg_b200 import ops
a_mn, b_mn, bn = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(0)
for (M, N, K) in [(128, 256, 64), (128, 256, 128), (384, 512, 256), (7552, 2304, 768)]:
    A = torch.randn(M, K, generator=g, device=dev).to(torch.bfloat16)
    B = torch.randn(N, K, generator=g, device=dev).to(torch.bfloat16)
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    out = torch.full((M, N), float('nan'), device=dev)
    ops.gemm(As, Bs, out, M=M, N=N, K=K, a_mn_major=bool(a_mn), b_mn_major=bool(b_mn), block_n=bn)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = (out - ref).abs().max().item()
    print(f'a_mn={a_mn} b_mn={b_mn} bn={bn} M={M} N={N} K={K} maxerr={err:.4g} nan={int(torch.isnan(out).sum())}', flush=True)
# timing on the c_attn shape
M, N, K = 7552, 2304, 768
A = torch.randn(M, K, device=dev).to(torch.bfloat16); B = torch.randn(N, K, device=dev).to(torch.bfloat16)
As = A.t().contiguous() if a_mn else A
Bs = B.t().contiguous() if b_mn else B
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(5): ops.gemm(As, Bs, out, M=M, N=N, K=K, a_mn_major=bool(a_mn), b_mn_major=bool(b_mn), block_n=bn)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.gemm(As, Bs, out, M=M, N=N, K=K, a_mn_major=bool(a_mn), b_mn_major=bool(b_mn), block_n=bn)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f'  timing {M}x{N}x{K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s', flush=True)
""" % ROOT

if __name__ == "__main__":
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            for bn in (128, 256):
                try:
                    r = subprocess.run([sys.executable, "-c", CHILD, str(a_mn), str(b_mn), str(bn)],
                                       capture_output=True, text=True, timeout=120)
                    print(r.stdout, end="")
                    if r.returncode != 0:
                        print(f"variant a_mn={a_mn} b_mn={b_mn} bn={bn} FAILED rc={r.returncode}\n{r.stderr[-1500:]}")
                except subprocess.TimeoutExpired:
                    print(f"variant a_mn={a_mn} b_mn={b_mn} bn={bn} TIMEOUT")
