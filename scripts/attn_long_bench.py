"""Attention backward at the extended lengths of configs[4]: tiled tcgen05 kernels vs the mma.sync pair."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmtg_b200 import ops
dev = torch.device("cuda:0")
NH = 12
for B, L in ((32, 436), (16, 636), (16, 1016)):
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(B * L, 3 * NH * 64, generator=g, device=dev).to(torch.bfloat16)
    mask = (torch.rand(B, L, generator=g, device=dev) > 0.2).to(torch.int32); mask[:, 0] = 1
    dout = (torch.randn(B * L, NH * 64, generator=g, device=dev) * 0.1).to(torch.bfloat16)
    out, lse = ops.attn_fwd(qkv, mask, B, L, NH, impl=2)
    res = {}
    for impl in (1, 2):
        fn = lambda: ops.attn_bwd(qkv, mask, out, dout, lse, B, L, NH, impl=impl)
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        res[impl] = e0.elapsed_time(e1) / 10 * 1e3
    print(f"B={B} L={L}: bwd mma.sync {res[1]:.1f} us, tcgen05 tiled {res[2]:.1f} us", flush=True)
