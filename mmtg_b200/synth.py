"""Deterministic synthetic inputs, token table and weights for MMTG (no datasets or checkpoints
are available offline — BASELINE.json prescribes random-init weights and synthetic embeddings).

Layout rules restate `MyDataset.__getitem__` / `convert_topic` / `convert_lyrics2ids`
(/root/reference/src/MyDataset.py:34-118); SURVEY.md §8(d) fixes the distributions and seeds.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from .configs import data_config as _data_config
from .configs import model_cfgs as _model_cfgs

PAD, START, EOS, UNK, CLS, SEP = 0, 1, 2, 100, 101, 102
VOCAB_SIZE = 13317
FIRST_REAL_ID = 104  # ids below are specials / unused slots of the BERT vocab


def make_token_table(seed: int = 4321, vocab_size: int = VOCAB_SIZE, dim: int = 2048) -> np.ndarray:
    """token id -> WenLan-like embedding, fp32 [V, dim], unit-norm rows
    (stands in for vocab/token_id2emb_dict.pkl, src/model.py:221-223)."""
    rng = np.random.default_rng(seed)
    t = rng.standard_normal((vocab_size, dim), dtype=np.float32)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    return t


def _unit_rows(rng, *shape):
    x = rng.standard_normal(shape).astype(np.float32)
    x /= np.linalg.norm(x, axis=-1, keepdims=True)
    return x


def make_batch(batch_size: int, seed: int = 1234, data_config=None, vocab_size: int = VOCAB_SIZE,
               ratings=None) -> "OrderedDict[str, np.ndarray]":
    """One collated batch with exactly the fields/dtypes/layout the reference DataLoader yields."""
    dc = data_config or _data_config()
    P, S, n_sent = dc.topic_prompt_length, dc.max_sent_length, dc.max_seq_length // (dc.max_sent_length + 2)
    D = dc.wenlan_emb_size
    rng = np.random.default_rng(seed)
    B = batch_size
    out = OrderedDict()
    topic_ids = np.zeros((B, P), np.int64)
    tpw_mask = np.zeros((B, P), np.int64)
    for b in range(B):
        n = int(rng.integers(6, P + 1))
        topic_ids[b, :n] = rng.integers(FIRST_REAL_ID, vocab_size, n)
        tpw_mask[b, :n] = 1
    out["topic_ids"] = topic_ids
    out["tpw_attention_mask"] = tpw_mask
    out["tpw_type_ids"] = tpw_mask.copy()  # 1 on real prompt tokens (MyDataset.py:69)
    out["topic_emb"] = _unit_rows(rng, B, D)
    out["img_embs"] = _unit_rows(rng, B, 5, D)
    out["r_embs"] = _unit_rows(rng, B, 5, D)
    T = n_sent * (S + 2) + 1
    targets = np.zeros((B, T), np.int64)
    mask = np.zeros((B, T), np.int64)
    types = np.zeros((B, T), np.int64)
    for b in range(B):
        pos = 0
        for s in range(n_sent):
            pair = s // 2
            tid = 1 if pair == 4 else pair + 1  # MyDataset.py:99-102
            n = int(rng.integers(5, S + 1))
            targets[b, pos] = START
            mask[b, pos] = 1
            targets[b, pos + 1:pos + 1 + n] = rng.integers(FIRST_REAL_ID, vocab_size, n)
            mask[b, pos + 1:pos + 1 + n] = 1
            types[b, pos + 1:pos + 1 + n] = tid
            targets[b, pos + S + 1] = EOS
            mask[b, pos + S + 1] = 1
            pos += S + 2
        targets[b, pos] = SEP
        mask[b, pos] = 1
    out["targets"] = targets
    out["attention_mask"] = mask
    out["type_ids"] = types
    if ratings is None:
        ratings = rng.integers(1, 6, B)
    out["rating"] = np.asarray(ratings, np.int64)
    return out


def batch_to_torch(batch, device="cpu"):
    return {k: torch.as_tensor(v).to(device) for k, v in batch.items()}


def state_dict_keys(n_layer: int = 12):
    """The 193 state_dict entries of the reference model, in registration order."""
    keys = ["encoder.topic_fc.weight", "encoder.topic_fc.bias"]
    for m in ("image", "text"):
        keys += [f"encoder.rnns_{m}.{n}" for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
    for i in (1, 2, 3):
        keys += [f"ln_layer{i}.weight", f"ln_layer{i}.bias"]
    for m in ("img", "text"):
        for n in ("query", "key", "value"):
            keys += [f"{m}_inner_atten_layer.{n}.weight", f"{m}_inner_atten_layer.{n}.bias"]
    for i in range(5):
        keys += [f"mm_atten_layer.att_matrices.{i}.weight", f"mm_atten_layer.att_matrices.{i}.bias"]
    keys += ["mm_atten_layer.out_linear.weight", "mm_atten_layer.out_linear.bias"]
    keys += ["decoder.projector_layer1.weight", "decoder.projector_layer1.bias",
             "decoder.projector_layer2.weight", "decoder.projector_layer2.bias"]
    g = "decoder.gpt2.transformer."
    keys += [g + "wte.weight", g + "wpe.weight"]
    for l in range(n_layer):
        h = f"{g}h.{l}."
        keys += [h + "ln_1.weight", h + "ln_1.bias", h + "attn.c_attn.weight", h + "attn.c_attn.bias",
                 h + "attn.c_proj.weight", h + "attn.c_proj.bias", h + "ln_2.weight", h + "ln_2.bias",
                 h + "mlp.c_fc.weight", h + "mlp.c_fc.bias", h + "mlp.c_proj.weight", h + "mlp.c_proj.bias"]
    keys += [g + "ln_f.weight", g + "ln_f.bias", "decoder.gpt2.lm_head.weight"]
    return keys


def make_state_dict(seed: int = 0, model_cfgs=None, gpt2_cfg=None, vocab_size: int = VOCAB_SIZE,
                    perturb: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """Random-init weights in the reference's state_dict layout (all fp32).

    Distributions follow the reference's init (xavier/orthogonal for the encoder,
    src/model.py:83-88; nn.Linear defaults; HF GPT-2 normal(0, 0.02) with 1/sqrt(2·n_layer) on
    residual projections). `perturb=True` additionally randomises biases and LayerNorm affine
    parameters so that parity tests cannot pass with those terms dropped.
    """
    cfg = model_cfgs or _model_cfgs
    g2 = dict(n_embd=768, n_head=12, n_layer=12, n_positions=1024, initializer_range=0.02)
    if gpt2_cfg:
        g2.update(gpt2_cfg)
    gen = torch.Generator().manual_seed(seed)
    H, Din, E, NL = cfg["topic"]["hidden_dim"], cfg["topic"]["input_dim"], g2["n_embd"], g2["n_layer"]

    def normal(*shape, std):
        return torch.randn(*shape, generator=gen) * std

    def uniform(*shape, bound):
        return (torch.rand(*shape, generator=gen) * 2 - 1) * bound

    def xavier(out_f, in_f):
        return normal(out_f, in_f, std=math.sqrt(2.0 / (in_f + out_f)))

    def orthogonal(rows, cols):
        a = torch.randn(rows, cols, generator=gen)
        q, r = torch.linalg.qr(a)
        return q * torch.sign(torch.diagonal(r)).unsqueeze(0)

    def lin_w(out_f, in_f):
        return uniform(out_f, in_f, bound=1.0 / math.sqrt(in_f))

    def lin_b(out_f, in_f):
        return uniform(out_f, bound=1.0 / math.sqrt(in_f))

    def ln_w(n):
        return 1.0 + (normal(n, std=0.05) if perturb else torch.zeros(n))

    def ln_b(n):
        return normal(n, std=0.05) if perturb else torch.zeros(n)

    def gbias(n):
        return normal(n, std=0.02) if perturb else torch.zeros(n)

    sd = OrderedDict()
    sd["encoder.topic_fc.weight"] = xavier(H, Din)
    sd["encoder.topic_fc.bias"] = lin_b(H, Din)
    for m in ("image", "text"):
        sd[f"encoder.rnns_{m}.weight_ih_l0"] = xavier(3 * H, Din)
        sd[f"encoder.rnns_{m}.weight_hh_l0"] = orthogonal(3 * H, H)
        sd[f"encoder.rnns_{m}.bias_ih_l0"] = uniform(3 * H, bound=1.0 / math.sqrt(H))
        sd[f"encoder.rnns_{m}.bias_hh_l0"] = uniform(3 * H, bound=1.0 / math.sqrt(H))
    for i in (1, 2, 3):
        sd[f"ln_layer{i}.weight"] = ln_w(H)
        sd[f"ln_layer{i}.bias"] = ln_b(H)
    for m in ("img", "text"):
        for n in ("query", "key", "value"):
            sd[f"{m}_inner_atten_layer.{n}.weight"] = lin_w(H, H)
            sd[f"{m}_inner_atten_layer.{n}.bias"] = lin_b(H, H)
    for i in range(cfg["seq_len"]):
        sd[f"mm_atten_layer.att_matrices.{i}.weight"] = lin_w(cfg["MM_ATT"]["attention_dim"], H)
        sd[f"mm_atten_layer.att_matrices.{i}.bias"] = lin_b(cfg["MM_ATT"]["attention_dim"], H)
    sd["mm_atten_layer.out_linear.weight"] = lin_w(2048, H)
    sd["mm_atten_layer.out_linear.bias"] = lin_b(2048, H)
    sd["decoder.projector_layer1.weight"] = lin_w(512, 2048)
    sd["decoder.projector_layer1.bias"] = lin_b(512, 2048)
    sd["decoder.projector_layer2.weight"] = lin_w(E, 512)
    sd["decoder.projector_layer2.bias"] = lin_b(E, 512)
    std = g2["initializer_range"]
    p = "decoder.gpt2.transformer."
    sd[p + "wte.weight"] = normal(vocab_size, E, std=std)
    sd[p + "wpe.weight"] = normal(g2["n_positions"], E, std=std)
    for l in range(NL):
        h = f"{p}h.{l}."
        sd[h + "ln_1.weight"] = ln_w(E)
        sd[h + "ln_1.bias"] = ln_b(E)
        sd[h + "attn.c_attn.weight"] = normal(E, 3 * E, std=std)
        sd[h + "attn.c_attn.bias"] = gbias(3 * E)
        sd[h + "attn.c_proj.weight"] = normal(E, E, std=std / math.sqrt(2 * NL))
        sd[h + "attn.c_proj.bias"] = gbias(E)
        sd[h + "ln_2.weight"] = ln_w(E)
        sd[h + "ln_2.bias"] = ln_b(E)
        sd[h + "mlp.c_fc.weight"] = normal(E, 4 * E, std=std)
        sd[h + "mlp.c_fc.bias"] = gbias(4 * E)
        sd[h + "mlp.c_proj.weight"] = normal(4 * E, E, std=std / math.sqrt(2 * NL))
        sd[h + "mlp.c_proj.bias"] = gbias(E)
    sd[p + "ln_f.weight"] = ln_w(E)
    sd[p + "ln_f.bias"] = ln_b(E)
    sd["decoder.gpt2.lm_head.weight"] = sd[p + "wte.weight"]  # tied
    assert list(sd.keys()) == state_dict_keys(NL)
    return sd
