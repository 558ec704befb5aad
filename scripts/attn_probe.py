import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmtg_b200 import ops
dev = torch.device("cuda:0")
for (B, L) in [(1, 64), (1, 128), (1, 236), (2, 436)]:
    NH = 12
    g = torch.Generator(device=dev).manual_seed(L)
    qkv = torch.randn(B * L, 3 * NH * 64, generator=g, device=dev).to(torch.bfloat16)
    mask = (torch.rand(B, L, generator=g, device=dev) > 0.25).to(torch.int32)
    mask[:, 0] = 1
    out, lse = ops.attn_fwd(qkv, mask, B, L, NH)
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(B, L, 3, NH, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / 8.0
    keep = torch.ones(L, L, dtype=torch.bool, device=dev).tril().view(1, 1, L, L) & (mask.view(B, 1, 1, L) != 0)
    s = s.masked_fill(~keep, float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * L, NH * 64)
    print(B, L, "max err", (out.float() - ref).abs().max().item(), "lse err", (lse - torch.logsumexp(s, -1)).abs().max().item(), flush=True)
