"""Row kernels + attention vs plain PyTorch fp32 references of the same ops (GPU)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M", [7552, 473])
def test_layernorm_bwd_split(cuda, M):
    """The engine's split of the decoder LayerNorm backward: row kernel on the chain (dx += ...,
    bf16 copy; no parameter gradients requested) + ln_param_grads on the side stream, against
    autograd of torch.nn.functional.layer_norm."""
    from mmtg_b200 import ops
    E = 768
    g = torch.Generator(device=cuda).manual_seed(M)
    x = torch.randn(M, E, generator=g, device=cuda) * 2 + 0.3
    gamma = 1 + 0.1 * torch.randn(E, generator=g, device=cuda)
    beta = 0.1 * torch.randn(E, generator=g, device=cuda)
    _y16, _y32, mean, rstd = ops.layernorm_fwd(x, gamma, beta, want_f32=True)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (E,), gr, br, 1e-5)
    dy = (torch.randn(M, E, generator=g, device=cuda) * 0.1).to(torch.bfloat16)
    ref.backward(dy.float())
    base = torch.randn(M, E, generator=g, device=cuda)
    dx = base.clone()
    dx16 = torch.empty(M, E, device=cuda, dtype=torch.bfloat16)
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, True, None, None, dx16, None)
    assert torch.allclose(dx - base, xr.grad, atol=2e-5, rtol=1e-4)
    assert torch.equal(dx16, dx.to(torch.bfloat16))
    dg, db, cs = (torch.zeros(E, device=cuda) for _ in range(3))
    ops.ln_param_grads(dy, x, mean, rstd, dx16, dg, db, cs)
    assert torch.allclose(dg, gr.grad, atol=1e-3 * math.sqrt(M), rtol=1e-4)
    assert torch.allclose(db, br.grad, atol=1e-3 * math.sqrt(M), rtol=1e-4)
    assert torch.allclose(cs, dx16.float().sum(0), atol=1e-3 * math.sqrt(M), rtol=1e-4)


@pytest.mark.parametrize("M,E", [(472, 768), (7552, 768), (160, 512), (3, 512)])
def test_layernorm_fwd_bwd(cuda, M, E):
    from mmtg_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(M + E)
    x = torch.randn(M, E, generator=g, device=cuda) * 2 + 0.3
    gamma = 1 + 0.1 * torch.randn(E, generator=g, device=cuda)
    beta = 0.1 * torch.randn(E, generator=g, device=cuda)
    y16, y32, mean, rstd = ops.layernorm_fwd(x, gamma, beta, want_f32=True)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (E,), gr, br, 1e-5)
    assert torch.allclose(y32, ref, atol=2e-5, rtol=1e-5)
    assert torch.allclose(y16.float(), ref, atol=2e-2, rtol=8e-3)
    for dt in (torch.float32, torch.bfloat16):
        dy = (torch.randn(M, E, generator=g, device=cuda) * 0.1).to(dt)
        for t in (xr, gr, br):
            t.grad = None
        ref.backward(dy.float(), retain_graph=True)
        base = torch.randn(M, E, generator=g, device=cuda)
        dx = base.clone()
        dg, db = torch.zeros(E, device=cuda), torch.zeros(E, device=cuda)
        dx16 = torch.empty(M, E, device=cuda, dtype=torch.bfloat16)
        cs = torch.zeros(E, device=cuda)
        ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, True, dg, db, dx16, cs)
        assert torch.allclose(dx - base, xr.grad, atol=2e-5, rtol=1e-4)
        assert torch.equal(dx16, dx.to(torch.bfloat16))
        assert torch.allclose(cs, dx.sum(0), atol=1e-3 * math.sqrt(M), rtol=1e-4)
        assert torch.allclose(dg, gr.grad, atol=1e-3 * math.sqrt(M), rtol=1e-4)
        assert torch.allclose(db, br.grad, atol=1e-3 * math.sqrt(M), rtol=1e-4)


def test_colsum(cuda):
    from mmtg_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(0)
    for M, N in [(7552, 768), (472, 2304), (160, 1536), (32, 512)]:
        x = torch.randn(M, N, generator=g, device=cuda)
        out = torch.ones(N, device=cuda)
        c16 = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
        ops.colsum(x, out, c16)
        assert torch.allclose(out, 1 + x.sum(0), atol=1e-3 * math.sqrt(M), rtol=1e-5)
        assert torch.equal(c16, x.to(torch.bfloat16))
        out = torch.zeros(N, device=cuda)
        ops.colsum(c16, out)
        assert torch.allclose(out, c16.float().sum(0), atol=1e-3 * math.sqrt(M), rtol=1e-5)


def _ref_attention(qkv, mask, B, L, NH):
    E = NH * 64
    q, k, v = qkv.float().view(B, L, 3, NH, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / 8.0
    keep = torch.ones(L, L, dtype=torch.bool, device=qkv.device).tril().view(1, 1, L, L) & (mask.view(B, 1, 1, L) != 0)
    s = s.masked_fill(~keep, float("-inf"))
    o = torch.softmax(s, -1) @ v
    return o.transpose(1, 2).reshape(B * L, E), torch.logsumexp(s, -1)


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("B,L", [(2, 236), (3, 64), (1, 17), (2, 436), (1, 1016), (2, 129), (1, 256), (2, 636), (1, 300), (1, 1024)])
def test_attention_fwd_bwd(cuda, B, L, impl):
    """bf16 tensor-core attention vs fp32 reference: out |Δ| <= 2e-2 (bf16 output rounding of
    O(1) values + bf16 P), gradients relative L2 error <= 2e-2."""
    from mmtg_b200 import ops
    NH = 12
    g = torch.Generator(device=cuda).manual_seed(B * 1000 + L)
    qkv = torch.randn(B * L, 3 * NH * 64, generator=g, device=cuda).to(torch.bfloat16)
    mask = (torch.rand(B, L, generator=g, device=cuda) > 0.25).to(torch.int32)
    mask[:, 0] = 1
    out, lse = ops.attn_fwd(qkv, mask, B, L, NH, impl=impl)
    qr = qkv.float().requires_grad_(True)
    ref, ref_lse = _ref_attention(qr, mask, B, L, NH)
    assert (out.float() - ref).abs().max().item() < 2e-2
    assert torch.allclose(lse, ref_lse, atol=2e-3, rtol=1e-4)
    dout = (torch.randn(B * L, NH * 64, generator=g, device=cuda) * 0.1).to(torch.bfloat16)
    ref.backward(dout.float())
    # impl 2: whole-head tcgen05 kernel for L <= 256, the tiled tcgen05 pair (dK/dV per key block, dQ per
    # query block) above
    dqkv = ops.attn_bwd(qkv, mask, out, dout, lse, B, L, NH, impl=impl)
    rel = (dqkv.float() - qr.grad).norm() / qr.grad.norm()
    assert rel.item() < 2e-2, rel.item()
    E = NH * 64
    for i, name in enumerate("qkv"):
        a, b = dqkv.float()[:, i * E:(i + 1) * E], qr.grad[:, i * E:(i + 1) * E]
        assert ((a - b).norm() / b.norm()).item() < 2e-2, name
