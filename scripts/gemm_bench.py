"""Times the tcgen05 GEMM on the decoder shapes (CUDA events, L2-flushed between launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmtg_b200 import ops

dev = torch.device("cuda:0")
only = sys.argv[1] if len(sys.argv) > 1 else None
shapes = [  # name, M, N, K, out dtype, a_mn, b_mn, split
    ("c_attn", 7552, 2304, 768, torch.bfloat16, 0, 0, 1),
    ("c_attn_f32out", 7552, 2304, 768, torch.float32, 0, 0, 1),
    ("attn_proj", 7552, 768, 768, torch.bfloat16, 0, 0, 1),
    ("c_fc", 7552, 3072, 768, torch.bfloat16, 0, 0, 1),
    ("mlp_proj", 7552, 768, 3072, torch.bfloat16, 0, 0, 1),
    ("lm_head", 7552, 13317, 768, torch.float32, 0, 0, 1),
    ("wgrad_fc", 768, 3072, 7552, torch.float32, 1, 1, 1),
    ("wgrad_proj", 768, 768, 7552, torch.float32, 1, 1, 4),
    ("tiny", 128, 256, 64, torch.bfloat16, 0, 0, 1),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, M, N, K, odt, a_mn, b_mn, split in shapes:
    if only and name != only:
        continue
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    out = torch.zeros(M, N, device=dev, dtype=odt)
    for bn in (128, 256):
        kw = dict(M=M, N=N, K=K, a_mn_major=bool(a_mn), b_mn_major=bool(b_mn), block_n=bn, split_k=split)
        for _ in range(3):
            ops.gemm(As, Bs, out, **kw)
        ts = []
        for _ in range(5 if not only else 1):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.gemm(As, Bs, out, **kw); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print(f"{name:14s} bn={bn} {M}x{N}x{K}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.0f} TFLOP/s", flush=True)
