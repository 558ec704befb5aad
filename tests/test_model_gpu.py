"""End-to-end parity of the CUDA path (through the drop-in Python surface, which calls the C-ABI
engine) against the CPU oracle and the committed reference goldens.

Tolerances (BASELINE.md §5, bf16 tensor-core GEMMs vs fp32 reference): logits max |Δ| <= 0.05,
mean |Δ| <= 0.01; HF loss / MyLoss |Δ| <= 2e-3 (relative for large values); KL |Δ| <= 5e-3;
gradients: per-tensor relative L2 error <= 5e-2 against fp32 autograd of the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def world(cuda):
    from mmtg_b200 import synth
    from mmtg_b200.configs import data_config, model_cfgs
    from mmtg_b200.model import MMTG
    table = synth.make_token_table()
    sd = synth.make_state_dict(0)
    model = MMTG(model_cfgs, data_config(), 13317, train_flag=True, token_table=table)
    model.set_dropout(0.0, 0.0, 0.0)  # parity with the reference is defined at p = 0 (test_dropout_gpu.py covers p > 0)
    model.load_state_dict(sd)
    model.to(cuda)
    return model, sd, table


def _batch(B, seed, cuda, ratings=None):
    from mmtg_b200 import synth
    host = synth.batch_to_torch(synth.make_batch(B, seed=seed, ratings=ratings))
    return host, {k: v.to(cuda) for k, v in host.items()}


def test_forward_vs_golden_and_oracle(world, cuda):
    from mmtg_b200.configs import data_config, model_cfgs
    from mmtg_b200.loss import MyLoss
    from oracle import mmtg_oracle as O
    model, sd, table = world
    g = np.load(os.path.join(G, "c1_forward_loss_b2.npz"))
    host, dev = _batch(2, 1234, cuda, ratings=np.array([5, 2]))
    with torch.no_grad():
        hf, kl, logits = model(dev)
        crit = MyLoss(data_config(), model_cfgs)
        my = [crit(logits, dev["targets"], dev["rating"], s).item() for s in (1, 2, 3)]
    assert logits.shape == (2, 236, 13317) and logits.dtype == torch.float32 and logits.is_contiguous()
    lg = logits.cpu()
    # reference goldens (executed reference, fp32)
    assert np.abs(lg[:, ::5, ::97].numpy() - g["logits_sub"]).max() <= 0.05
    assert np.abs(lg[:, [0, 14, 15, 100, 235], :].numpy() - g["logits_rows"]).max() <= 0.05
    assert abs(hf.item() - float(g["hf_loss"])) <= 2e-3
    assert abs(kl.item() - float(g["kl"])) <= 5e-3
    for i, s in enumerate((1, 2, 3)):
        ref = float(g[f"myloss_stage{s}"])
        assert abs(my[i] - ref) <= 2e-3 * max(1.0, abs(ref)), (s, my[i], ref)
    # full-tensor check against the oracle
    with torch.no_grad():
        ohf, okl, ologits = O.mmtg_forward(sd, torch.from_numpy(table), host, data_config(), True)
    diff = (lg - ologits).abs()
    assert diff.max().item() <= 0.05 and diff.mean().item() <= 0.01, (diff.max().item(), diff.mean().item())
    agree = (lg.argmax(-1) == ologits.argmax(-1)).float().mean().item()
    assert agree > 0.9, agree


def test_gradients_vs_golden(world, cuda):
    """Restated train step (src/train.py:188-193): total = MyLoss.mean() + 0.2 * kl.mean()."""
    from mmtg_b200.configs import data_config, model_cfgs
    from mmtg_b200.loss import MyLoss
    model, sd, table = world
    g = np.load(os.path.join(G, "c1_forward_loss_b2.npz"))
    host, dev = _batch(2, 1234, cuda, ratings=np.array([5, 2]))
    crit = MyLoss(data_config(), model_cfgs)
    model.zero_grad(set_to_none=True)
    hf, kl, logits = model(dev)
    loss = crit(logits, dev["targets"], dev["rating"], 3)
    total = loss.mean() + 0.2 * kl.mean()
    total.backward()
    assert abs(total.item() - float(g["total_loss"])) <= 3e-3
    names = [str(n) for n in g["grad_names"]]
    params = dict(model.named_parameters())
    bad = []
    for i, n in enumerate(names):
        gr = params[n].grad.detach().float().flatten().cpu()
        ref_norm = float(g["grad_norms"][i])
        idx = torch.linspace(0, gr.numel() - 1, 32).long()
        err_norm = abs(gr.norm().item() - ref_norm)
        err_samp = np.abs(gr[idx].numpy() - g["grad_samples"][i]).max()
        # absolute floors: key-bias / gate-bias gradients are analytically zero; bf16 rounding of the
        # dqkv operand leaves <= 1e-4 of noise there (reference value ~1e-9)
        if err_norm > 5e-2 * ref_norm + 1e-4 or err_samp > 0.15 * np.abs(g["grad_samples"][i]).max() + 1e-5:
            bad.append((n, err_norm, ref_norm, err_samp))
    assert not bad, bad


def test_gradients_full_vs_oracle_and_generic_loss(world, cuda):
    """Per-tensor relative L2 error of every gradient vs fp32 autograd over the oracle, using the
    GENERIC loss path (dense fp32 dlogits) — must agree with the fused path too."""
    from mmtg_b200.configs import data_config, model_cfgs
    from mmtg_b200.loss import MyLoss
    from oracle import mmtg_oracle as O
    model, sd, table = world
    host, dev = _batch(3, 77, cuda)
    crit = MyLoss(data_config(), model_cfgs)
    # oracle gradients
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "decoder.gpt2.lm_head.weight"}
    params["decoder.gpt2.lm_head.weight"] = params["decoder.gpt2.transformer.wte.weight"]
    ohf, okl, ologits = O.mmtg_forward(params, torch.from_numpy(table), host, data_config(), True)
    ototal = O.my_loss(ologits, host["targets"], host["rating"], 2).mean() + 0.2 * okl.mean()
    ototal.backward()
    results = {}
    for mode in ("fused", "generic"):
        model.zero_grad(set_to_none=True)
        hf, kl, logits = model(dev)
        lg = logits if mode == "fused" else logits * 1.0  # the multiply hides the fused tag
        total = crit(lg, dev["targets"], dev["rating"], 2).mean() + 0.2 * kl.mean()
        total.backward()
        assert abs(total.item() - ototal.item()) <= 3e-3
        results[mode] = {n: p.grad.detach().float().cpu().clone() for n, p in model.named_parameters()}
    bad = []
    for n, ref in ((k, v.grad) for k, v in params.items() if k != "decoder.gpt2.lm_head.weight"):
        for mode in results:
            # key-bias / gate-bias gradients are analytically zero (softmax shift invariance):
            # an absolute floor keeps the relative test meaningful for them
            err = (results[mode][n] - ref).norm().item()
            if err > 5e-2 * ref.norm().item() + 1e-4:
                bad.append((mode, n, err, ref.norm().item()))
    assert not bad, bad


def test_state_dict_roundtrip_and_eval_no_grad(world, cuda):
    from mmtg_b200 import synth
    model, sd, table = world
    assert list(model.state_dict().keys()) == synth.state_dict_keys()
    out = model.state_dict()
    for k, v in sd.items():
        assert torch.equal(out[k].cpu(), v), k
    # DataParallel-style prefix + transformers-4.x mask buffers are tolerated
    sd2 = {"module." + k: v for k, v in sd.items()}
    sd2["module.decoder.gpt2.transformer.h.0.attn.bias"] = torch.ones(1, 1, 4, 4)
    sd2["module.decoder.gpt2.transformer.h.0.attn.masked_bias"] = torch.tensor(-1e4)
    model.load_state_dict(sd2)


def test_fused_train_step_and_graph_match_autograd_path(world, cuda):
    """MMTG.fused_train_step (autograd-free driver) and its CUDA-graph replay produce the same
    loss and gradients as forward() + MyLoss + backward()."""
    from mmtg_b200.configs import data_config, model_cfgs
    from mmtg_b200.loss import MyLoss
    model, sd, table = world
    host, dev = _batch(2, 1234, cuda, ratings=np.array([5, 2]))
    crit = MyLoss(data_config(), model_cfgs)
    model.zero_grad(set_to_none=True)
    hf, kl, logits = model(dev)
    total = crit(logits, dev["targets"], dev["rating"], 3).mean() + 0.2 * kl.mean()
    total.backward()
    ref = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    t2, l2, k2 = model.fused_train_step(dev, 3, 0.2)
    assert abs(t2.item() - total.item()) < 1e-5
    for n, p in model.named_parameters():
        err = (p.grad - ref[n]).norm().item()
        assert err <= 2e-3 * ref[n].norm().item() + 1e-6, (n, err)  # split-K atomics reorder sums


@pytest.mark.parametrize("max_sent_length", [40])
def test_extended_lyrics_length_forward_backward(cuda, max_sent_length):
    """BASELINE.json configs[4]: extended lyrics length (L = 15 + 10*(msl+2) + 1 = 436) — forward
    and gradient parity vs the oracle with a 2-layer decoder (keeps the CPU oracle fast)."""
    from mmtg_b200 import synth
    from mmtg_b200.configs import data_config, model_cfgs
    from mmtg_b200.loss import MyLoss
    from mmtg_b200.model import MMTG
    from oracle import mmtg_oracle as O
    dc = data_config(max_sent_length=max_sent_length)
    g2 = {"n_layer": 2}
    table = synth.make_token_table()
    sd = synth.make_state_dict(1, gpt2_cfg=g2)
    model = MMTG(model_cfgs, dc, 13317, train_flag=True, token_table=table, gpt2_config=g2)
    model.set_dropout(0.0, 0.0, 0.0)
    model.load_state_dict(sd)
    model.to(cuda)
    host = synth.batch_to_torch(synth.make_batch(2, seed=11, data_config=dc, ratings=np.array([1, 5])))
    dev = {k: v.to(cuda) for k, v in host.items()}
    assert host["targets"].shape[1] == 10 * (max_sent_length + 2) + 1
    crit = MyLoss(dc, model_cfgs)
    hf, kl, logits = model(dev)
    total = crit(logits, dev["targets"], dev["rating"], 3).mean() + 0.2 * kl.mean()
    total.backward()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "decoder.gpt2.lm_head.weight"}
    params["decoder.gpt2.lm_head.weight"] = params["decoder.gpt2.transformer.wte.weight"]
    ctx, okl = O.fused_context(params, host)
    t_emb, i_emb = O.decoder_embed(torch.from_numpy(table), ctx, host["topic_ids"], host["targets"], 2 * (max_sent_length + 2))
    emb = torch.cat([t_emb, i_emb], 1)
    types = torch.cat([host["tpw_type_ids"], host["type_ids"]], 1)
    mask = torch.cat([host["tpw_attention_mask"], host["attention_mask"]], 1)
    h1 = torch.tanh(emb @ params["decoder.projector_layer1.weight"].t() + params["decoder.projector_layer1.bias"])
    x = h1 @ params["decoder.projector_layer2.weight"].t() + params["decoder.projector_layer2.bias"]
    ologits = O.gpt2_forward(params, x, types, mask, n_layer=2)
    ototal = O.my_loss(ologits, host["targets"], host["rating"], 3).mean() + 0.2 * okl
    ototal.backward()
    d = (logits.detach().cpu() - ologits.detach()).abs()
    assert d.max().item() <= 0.05 and d.mean().item() <= 0.01
    assert abs(total.item() - ototal.item()) <= 3e-3 * max(1.0, abs(ototal.item()))
    bad = []
    for n, p in model.named_parameters():
        ref = params[n].grad
        err = (p.grad.detach().cpu() - ref).norm().item()
        if err > 5e-2 * ref.norm().item() + 1e-4:
            bad.append((n, err, ref.norm().item()))
    assert not bad, bad
