"""Times attention fwd/bwd implementations at the train shape (B=32, L=236, 12 heads)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmtg_b200 import ops
dev = torch.device("cuda:0")
B, L, NH = 32, 236, 12
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(B * L, 3 * NH * 64, generator=g, device=dev).to(torch.bfloat16)
mask = (torch.rand(B, L, generator=g, device=dev) > 0.2).to(torch.int32); mask[:, 0] = 1
dout = (torch.randn(B * L, NH * 64, generator=g, device=dev) * 0.1).to(torch.bfloat16)
for impl in (1, 2):
    out, lse = ops.attn_fwd(qkv, mask, B, L, NH, impl=impl)
    dq = ops.attn_bwd(qkv, mask, out, dout, lse, B, L, NH, impl=impl)
    for name, fn in (("fwd", lambda: ops.attn_fwd(qkv, mask, B, L, NH, impl=impl)),
                     ("bwd", lambda: ops.attn_bwd(qkv, mask, out, dout, lse, B, L, NH, impl=impl))):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"impl={impl} {name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
