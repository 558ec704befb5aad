"""Pins the CPU oracle (oracle/mmtg_oracle.py) against golden vectors produced by executing the
unmodified reference (scripts/make_golden.py -> tests/golden/*.npz). fp32 vs fp32: the only
differences are accumulation order, so tolerances are tight (logits 2e-5 abs)."""
import os

import numpy as np
import pytest
import torch

from mmtg_b200 import synth
from mmtg_b200.configs import data_config
from oracle import mmtg_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def world():
    torch.set_num_threads(8)
    sd = synth.make_state_dict(0)
    table = torch.from_numpy(synth.make_token_table())
    return sd, table


def test_c1_forward_and_losses(world):
    sd, table = world
    g = np.load(os.path.join(G, "c1_forward_loss_b2.npz"))
    batch = synth.batch_to_torch(synth.make_batch(2, seed=1234, ratings=np.array([5, 2])))
    with torch.no_grad():
        hf, kl, logits = O.mmtg_forward(sd, table, batch, data_config(), True)
    assert logits.shape == (2, 236, 13317)
    assert abs(hf.item() - float(g["hf_loss"])) < 1e-5
    assert abs(kl.item() - float(g["kl"])) < 1e-5
    assert np.abs(logits[:, ::5, ::97].numpy() - g["logits_sub"]).max() < 2e-5
    assert np.abs(logits[:, [0, 14, 15, 100, 235], :].numpy() - g["logits_rows"]).max() < 2e-5
    assert np.abs(logits.sum(-1).numpy() - g["logits_rowsum"]).max() < 2e-3
    for stage in (1, 2, 3):
        v = O.my_loss(logits, batch["targets"], batch["rating"], stage).item()
        assert abs(v - float(g[f"myloss_stage{stage}"])) < 1e-5 * max(1.0, abs(v))


def test_c1_gradients(world):
    """autograd over the oracle == autograd over the reference (restated src/train.py:188-193)."""
    sd, table = world
    g = np.load(os.path.join(G, "c1_forward_loss_b2.npz"))
    batch = synth.batch_to_torch(synth.make_batch(2, seed=1234, ratings=np.array([5, 2])))
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "decoder.gpt2.lm_head.weight"}
    params["decoder.gpt2.lm_head.weight"] = params["decoder.gpt2.transformer.wte.weight"]
    hf, kl, logits = O.mmtg_forward(params, table, batch, data_config(), True)
    total = O.my_loss(logits, batch["targets"], batch["rating"], 3).mean() + 0.2 * kl.mean()
    total.backward()
    assert abs(total.item() - float(g["total_loss"])) < 1e-5
    names = [str(n) for n in g["grad_names"]]
    assert len(names) == 192
    for i, n in enumerate(names):
        gr = params[n].grad.flatten()
        ref_norm = float(g["grad_norms"][i])
        assert abs(gr.norm().item() - ref_norm) <= 2e-4 * ref_norm + 1e-7, n
        idx = torch.linspace(0, gr.numel() - 1, 32).long()
        assert np.abs(gr[idx].numpy() - g["grad_samples"][i]).max() <= 2e-4 * ref_norm + 1e-7, n


def test_loss_sweep(world):
    sd, table = world
    g = np.load(os.path.join(G, "loss_sweep_b4.npz"))
    b4 = synth.batch_to_torch(synth.make_batch(4, seed=77))
    with torch.no_grad():
        _, kl, logits = O.mmtg_forward(sd, table, b4, data_config(), True)
    assert abs(kl.item() - float(g["kl"])) < 1e-5
    assert np.abs(logits.sum(-1).numpy() - g["logits_rowsum"]).max() < 2e-3
    for key in g.files:
        if not key.startswith("r"):
            continue
        r = [int(c) for c in key[1:5]]
        stage = int(key[-1])
        v = O.my_loss(logits, b4["targets"], torch.tensor(r), stage).item()
        assert abs(v - float(g[key])) < 1e-5 * max(1.0, abs(v)), key


def test_generation(world):
    sd, table = world
    g = np.load(os.path.join(G, "generate_b1.npz"))
    one = synth.make_batch(1, seed=99)
    start = {k: v[0] for k, v in one.items() if k != "rating"}
    start["targets"] = np.asarray([1])
    ids, step_logits = O.sample_sequence(sd, table, dict(start), int(g["length"]), data_config(),
                                         temperature=1.0, top_k=1, top_p=0.0, repitition_penalty=1.0,
                                         return_logits=True)
    assert ids == g["greedy_ids"].tolist()
    sl = torch.stack(step_logits).numpy()
    assert np.abs(sl[:, ::13] - g["greedy_step_logits_sub"]).max() < 2e-5
    torch.manual_seed(5)
    ids2 = O.sample_sequence(sd, table, dict(start), 24, data_config(), temperature=1.1, top_k=10,
                             top_p=0.7, repitition_penalty=1.5)
    assert ids2 == g["sampled_ids_seed5"].tolist()


def test_top_k_top_p_known_answers():
    g = np.load(os.path.join(G, "generate_b1.npz"))
    for i in range(5):
        k, p = g[f"filt_kp_{i}"]
        y = O.top_k_top_p_filtering(torch.from_numpy(g[f"filt_in_{i}"].copy()), top_k=int(k), top_p=float(p))
        assert torch.isfinite(y).nonzero().flatten().tolist() == g[f"filt_keep_{i}"].tolist()
