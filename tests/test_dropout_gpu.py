"""GPT-2 dropout (embd / attention / resid, p = 0.1 in the reference's training forward) on the
CUDA path. RNG streams cannot match torch's, so parity is checked by INJECTING the kernels' own
counter-based masks (mmtg_dropout_mask) into the CPU oracle: with identical masks, logits,
losses and every gradient must agree to the same tolerances as the p = 0 tests."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

EMBD_SITE = 0xFFFF


def _mask(seed_t, site, p, n, cuda):
    from mmtg_b200 import _lib
    out = torch.empty(n, dtype=torch.uint8, device=cuda)
    _lib.check(_lib.lib().mmtg_dropout_mask(C.c_void_p(seed_t.data_ptr()), C.c_uint32(site), C.c_float(p),
                                            C.c_int64(n), C.c_void_p(out.data_ptr()), C.c_void_p(_lib.stream_ptr())),
               "mmtg_dropout_mask")
    return out


def test_mask_statistics_and_streams(cuda):
    seed = torch.tensor([1234], dtype=torch.int64, device=cuda)
    n = 1 << 22
    a = _mask(seed, 5, 0.1, n, cuda).float()
    b = _mask(seed, 6, 0.1, n, cuda).float()
    seed2 = torch.tensor([1235], dtype=torch.int64, device=cuda)
    c = _mask(seed2, 5, 0.1, n, cuda).float()
    for m in (a, b, c):
        assert abs(m.mean().item() - 0.9) < 2e-3
    # neighbouring elements share one hash: their keep flags must still be independent
    pair = (a[0::2] * a[1::2]).mean().item()
    assert abs(pair - 0.81) < 3e-3, pair
    for x, y in ((a, b), (a, c)):
        both = (x * y).mean().item()
        assert abs(both - 0.81) < 3e-3, both  # independent streams
    # shifted-copy check: site streams must not be translations of each other
    for sh in (1, 2, 64, 4096):
        assert abs((a[sh:] * b[:-sh]).mean().item() - 0.81) < 3e-3
    assert _mask(seed, 5, 0.0, 1024, cuda).min().item() == 1
    assert abs(_mask(seed, 9, 0.5, n, cuda).float().mean().item() - 0.5) < 2e-3


@pytest.mark.parametrize("L,impl", [(236, 2), (100, 2), (128, 2), (236, 1), (300, 1), (77, 1), (300, 0)])
def test_attention_dropout_matches_masked_reference(cuda, L, impl):
    """impl 2 = tcgen05 kernels, 1 = mma.sync kernels (the backward for L > 256), 0 = default dispatch."""
    from mmtg_b200 import _lib
    lib = _lib.lib()
    B, NH, E, p, site = 2, 12, 768, 0.1, 4 * 3
    g = torch.Generator().manual_seed(L)
    qkv = (torch.randn(B * L, 3 * E, generator=g) * 0.7).to(cuda).bfloat16()
    dout = (torch.randn(B * L, E, generator=g) * 0.5).to(cuda).bfloat16()
    kmask = torch.ones(B, L, dtype=torch.int32)
    kmask[1, L - 17:] = 0
    kmask = kmask.to(cuda)
    seed = torch.tensor([99], dtype=torch.int64, device=cuda)
    out = torch.empty(B * L, E, dtype=torch.bfloat16, device=cuda)
    lse = torch.empty(B, NH, L, device=cuda)
    delta = torch.empty(B, NH, L, device=cuda)
    dqkv = torch.zeros(B * L, 3 * E, dtype=torch.bfloat16, device=cuda)
    st = C.c_void_p(_lib.stream_ptr())
    vp = C.c_void_p
    _lib.check(lib.mmtg_attn_fwd_drop(vp(qkv.data_ptr()), vp(kmask.data_ptr()), vp(out.data_ptr()), vp(lse.data_ptr()),
                                      B, L, NH, vp(seed.data_ptr()), C.c_uint32(site), C.c_float(p), impl, st), "attn_fwd_drop")
    _lib.check(lib.mmtg_attn_bwd_drop(vp(qkv.data_ptr()), vp(kmask.data_ptr()), vp(out.data_ptr()), vp(dout.data_ptr()),
                                      vp(lse.data_ptr()), vp(delta.data_ptr()), vp(dqkv.data_ptr()), B, L, NH,
                                      vp(seed.data_ptr()), C.c_uint32(site), C.c_float(p), impl, st), "attn_bwd_drop")
    Lp = (L + 1) // 2 * 2
    keep = _mask(seed, site, p, B * NH * L * Lp, cuda).view(B, NH, L, Lp)[..., :L].float()
    x = qkv.float().view(B, L, 3, NH, 64).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    q, k, v = x[0], x[1], x[2]
    s = q @ k.transpose(-1, -2) / 8.0
    allow = torch.ones(L, L, dtype=torch.bool, device=cuda).tril().view(1, 1, L, L) & (kmask.view(B, 1, 1, L) != 0)
    w = torch.softmax(s.masked_fill(~allow, float("-inf")), -1) * keep / (1 - p)
    ref = (w @ v).transpose(1, 2).reshape(B * L, E)
    ref.backward(dout.float())
    dref = x.grad.permute(1, 3, 0, 2, 4).reshape(B * L, 3 * E)
    assert (out.float() - ref).abs().max().item() <= 0.03
    err = (dqkv.float() - dref).norm().item() / dref.norm().item()
    assert err <= 2e-2, err
    # and the masks actually bite: p = 0 gives a different output
    out0 = torch.empty_like(out)
    _lib.check(lib.mmtg_attn_fwd_drop(vp(qkv.data_ptr()), vp(kmask.data_ptr()), vp(out0.data_ptr()), vp(lse.data_ptr()),
                                      B, L, NH, vp(seed.data_ptr()), C.c_uint32(site), C.c_float(0.0), impl, st), "attn_fwd_drop")
    assert (out0.float() - out.float()).abs().max().item() > 0.05


def _site_masks(model, d, cuda):
    seed = model._drop_seed
    pe, pr, pa = model._drop_p
    B, L, E, NH, NL = d.B, d.L, d.E, d.NH, d.NL
    Lp = (L + 1) // 2 * 2
    m = {"p": (pe, pr, pa), "attn": [], "resid1": [], "resid2": []}
    m["embd"] = _mask(seed, EMBD_SITE, pe, B * L * E, cuda).view(B, L, E).cpu()
    for l in range(NL):
        m["attn"].append(_mask(seed, 4 * l, pa, B * NH * L * Lp, cuda).view(B, NH, L, Lp)[..., :L].cpu())
        m["resid1"].append(_mask(seed, 4 * l + 1, pr, B * L * E, cuda).view(B, L, E).cpu())
        m["resid2"].append(_mask(seed, 4 * l + 2, pr, B * L * E, cuda).view(B, L, E).cpu())
    return m


@pytest.mark.parametrize("n_layer,max_sent_length", [(12, None), (2, 30)])
def test_train_step_with_dropout_matches_masked_oracle(cuda, n_layer, max_sent_length):
    """(12, default): the benchmark configuration (L = 236, tcgen05 attention both ways);
    (2, 30): extended lyrics length (L = 336): the backward runs the mma.sync attention kernels."""
    from mmtg_b200 import synth
    from mmtg_b200.configs import data_config as _dc, model_cfgs
    from mmtg_b200.loss import MyLoss
    from mmtg_b200.model import MMTG
    from oracle import mmtg_oracle as O
    table = synth.make_token_table()
    g2 = {"n_layer": n_layer}
    dcfg = _dc() if max_sent_length is None else _dc(max_sent_length=max_sent_length)
    data_config = lambda: dcfg
    sd = synth.make_state_dict(0, gpt2_cfg=g2)
    model = MMTG(model_cfgs, dcfg, 13317, train_flag=True, token_table=table, gpt2_config=g2)
    model.load_state_dict(sd)
    model.to(cuda)
    assert model._drop_p == (0.1, 0.1, 0.1) and model.dropout_active()
    model.set_dropout_seed(4242)
    host = synth.batch_to_torch(synth.make_batch(2, seed=31, data_config=dcfg, ratings=np.array([4, 1])))
    dev = {k: v.to(cuda) for k, v in host.items()}
    crit = MyLoss(data_config(), model_cfgs)
    model.zero_grad(set_to_none=True)
    hf, kl, logits = model(dev)
    total = crit(logits, dev["targets"], dev["rating"], 2).mean() + 0.2 * kl.mean()
    total.backward()
    torch.cuda.synchronize()
    d = logits._mmtg_step.dims
    masks = _site_masks(model, d, cuda)  # the seed this step used is still in the device tensor
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "decoder.gpt2.lm_head.weight"}
    params["decoder.gpt2.lm_head.weight"] = params["decoder.gpt2.transformer.wte.weight"]
    ohf, okl, ologits = O.mmtg_forward(params, torch.from_numpy(table), host, data_config(), True, dropout=masks)
    ototal = O.my_loss(ologits, host["targets"], host["rating"], 2).mean() + 0.2 * okl.mean()
    ototal.backward()
    diff = (logits.detach().cpu() - ologits.detach()).abs()
    assert diff.max().item() <= 0.05 and diff.mean().item() <= 0.01, (diff.max().item(), diff.mean().item())
    assert abs(hf.item() - ohf.item()) <= 2e-3 and abs(total.item() - ototal.item()) <= 3e-3
    # dropout must have changed the result: the eval-mode oracle is far away
    with torch.no_grad():
        _, _, elogits = O.mmtg_forward(sd, torch.from_numpy(table), host, data_config(), True)
    assert (logits.detach().cpu() - elogits).abs().max().item() > 0.2
    bad = []
    for n, p in model.named_parameters():
        if n == "decoder.gpt2.lm_head.weight":
            continue
        ref = params[n].grad
        gr = p.grad.detach().float().cpu()
        err = (gr - ref).norm().item()
        if err > 5e-2 * ref.norm().item() + 1e-4:
            bad.append((n, err, ref.norm().item()))
    assert not bad, bad
    # a second step draws new masks; eval() switches dropout off
    hf2, _, logits2 = model(dev)
    assert (logits2.detach() - logits.detach()).abs().max().item() > 0.05
    model.eval()
    with torch.no_grad():
        _, _, l3 = model(dev)
    model.train()
    assert (l3.cpu() - elogits).abs().max().item() <= 0.05


def test_fused_step_and_graph_replay_advance_masks(cuda):
    """The autograd-free step (CUDA-graph body) carries the seed bump: replays draw new masks."""
    from mmtg_b200 import synth
    from mmtg_b200.configs import data_config, model_cfgs
    from mmtg_b200.model import MMTG
    model = MMTG(model_cfgs, data_config(), 13317, train_flag=True, token_table=synth.make_token_table())
    model.load_state_dict(synth.make_state_dict(0))
    model.to(cuda)
    dev = {k: v.to(cuda) for k, v in synth.batch_to_torch(synth.make_batch(2, seed=5)).items()}
    s0 = None
    losses = []
    for _ in range(3):
        for p in model.parameters():
            p.grad = None
        total, loss, kl = model.fused_train_step(dev, 2)
        losses.append(loss.item())
        s = int(model._drop_seed.item())
        assert s != s0
        s0 = s
    assert len(set(losses)) == 3 and all(np.isfinite(losses)), losses
