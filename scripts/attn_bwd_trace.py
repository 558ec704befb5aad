"""Stage timeline (globaltimer, CTA 0) of the tcgen05 attention backward at the train shape."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mmtg_b200 import _lib, ops
dev = torch.device("cuda:0")
B, L, NH = 32, 236, 12
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(B * L, 3 * NH * 64, generator=g, device=dev).to(torch.bfloat16)
mask = (torch.rand(B, L, generator=g, device=dev) > 0.2).to(torch.int32); mask[:, 0] = 1
dout = (torch.randn(B * L, NH * 64, generator=g, device=dev) * 0.1).to(torch.bfloat16)
out, lse = ops.attn_fwd(qkv, mask, B, L, NH, impl=2)
buf = torch.zeros(64, dtype=torch.int64, device=dev)
lib = _lib.lib()
lib.mmtg_attn_set_clk.argtypes = [C.c_void_p]
for _ in range(3):
    ops.attn_bwd(qkv, mask, out, dout, lse, B, L, NH, impl=2)
lib.mmtg_attn_set_clk(C.c_void_p(buf.data_ptr()))
ops.attn_bwd(qkv, mask, out, dout, lse, B, L, NH, impl=2)
torch.cuda.synchronize()
lib.mmtg_attn_set_clk(C.c_void_p(0))
t = buf.cpu().numpy().astype(np.float64)
t0 = t[32]
def rel(i): return round((t[i] - t0) / 1e3, 2) if t[i] > 0 else None
print("softmax warp 0: start 0, after prologue sync", rel(33))
print("control: entered", rel(0), "tma issued", rel(1), "block0 landed", rel(2))
for pr in range(3):
    print(f" pair {pr}: ctrl sdp issued {rel(3+pr*3)} pds ready {rel(4+pr*3)} mma2 issued {rel(5+pr*3)} | sm: enter {rel(34+pr*5)} sdp ready {rel(35+pr*5)} smem free {rel(36+pr*5)} tmem loaded {rel(37+pr*5)} arrived {rel(38+pr*5)}")
print("drain j0", rel(50), "drain j1", rel(52), "dq done", rel(60), "end", rel(61))
