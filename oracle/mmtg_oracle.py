"""ORACLE — test infrastructure only. Never imported by the product path (mmtg_b200/).

A CPU restatement, in plain PyTorch fp32 ops, of the reference algorithm for the MMTG hot path:
`MMTG.forward` (/root/reference/src/model.py:356-400), the HF GPT-2 arithmetic it calls
(transformers 5.5.0 models/gpt2/modeling_gpt2.py — third-party, not vendored by the reference,
pinned there as transformers==4.12.3 in requirements.txt:2; not installable offline),
`MyLoss.forward` (src/loss.py:45-74) and `sample_sequence` / `top_k_top_p_filtering`
(src/generate.py:64-145).

Parity pin: the reference has no tests, golden vectors or fixtures, so this file
is pinned against OUTPUTS OF THE REFERENCE ITSELF executed in the build container
(oracle/ref_import.py imports /root/reference/src unmodified; scripts/make_golden.py writes
tests/golden/*.npz; tests/test_oracle_golden.py checks this restatement against them, and
tests/test_oracle_vs_reference.py re-checks live whenever /root/reference is present).
Floating point, fp32 everywhere; tolerance vs the reference: 2e-5 abs on logits.

All functions are functional over a state_dict (name -> fp32 tensor) in the reference's layout,
and differentiable, so autograd on this file is also the gradient oracle.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-5


# --------------------------------------------------------------------------------------------
# encoder side
# --------------------------------------------------------------------------------------------
def gru_forward(x, w_ih, w_hh, b_ih, b_hh):
    """nn.GRU, 1 layer, h0 = 0, gate order (r, z, n). x: [S, B, D] -> [S, B, H].
    Reference: src/model.py:47-48,78-79 (torch.nn.GRU semantics)."""
    S, B, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    outs = []
    gi_all = x @ w_ih.t() + b_ih
    for t in range(S):
        gi = gi_all[t]
        gh = h @ w_hh.t() + b_hh
        i_r, i_z, i_n = gi.chunk(3, -1)
        h_r, h_z, h_n = gh.chunk(3, -1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h = (1 - z) * n + z * h
        outs.append(h)
    return torch.stack(outs, 0)


def encoder_forward(sd, topic_emb, img_embs, r_embs):
    """MultiModalEncoder.forward, src/model.py:63-81. Inputs [B,D], [S,B,D], [S,B,D]."""
    topic = (topic_emb @ sd["encoder.topic_fc.weight"].t() + sd["encoder.topic_fc.bias"]).unsqueeze(0)
    img = gru_forward(img_embs, *[sd[f"encoder.rnns_image.{n}"] for n in
                                  ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")])
    txt = gru_forward(r_embs, *[sd[f"encoder.rnns_text.{n}"] for n in
                                ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")])
    return topic, img, txt


def gaussian_priors(n: int = 5):
    """Discretised N(i, 1) over {0..n-1}, normalised. src/model.py:116-120 (scipy.stats.norm.pdf)."""
    idx = np.arange(n, dtype=np.float64)
    rows = []
    for i in range(n):
        pdf = np.exp(-0.5 * (idx - i) ** 2) / math.sqrt(2 * math.pi)
        rows.append(torch.tensor([v / pdf.sum() for v in pdf], dtype=torch.float32))
    return torch.stack(rows, 0)  # [query i, key j]


def alpha_attention(sd, prefix, x, heads: int = 4):
    """InnerModalAttentionLayer.forward, src/model.py:133-161. x: [B, S, H] -> ([B, S, H], kl)."""
    B, S, Hd = x.shape
    dh = Hd // heads

    def proj(n):
        y = x @ sd[f"{prefix}.{n}.weight"].t() + sd[f"{prefix}.{n}.bias"]
        return y.view(B, S, heads, dh).permute(0, 2, 1, 3)

    q, k, v = proj("query"), proj("key"), proj("value")
    probs = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), -1)  # [B, h, S, S]
    prior = gaussian_priors(S).to(x.device)
    kls = []
    for i in range(S):
        t = prior[i].view(1, 1, S).expand(B, heads, S)
        # KLDivLoss(reduction='batchmean'): sum(t * (log t - input)) / input.size(0)
        kls.append((t * (t.log() - probs[:, :, i, :].log())).sum() / B)
    ctx = (probs @ v).permute(0, 2, 1, 3).reshape(B, S, Hd)
    return ctx, torch.stack(kls).mean()


def beta_attention(sd, topic, img, txt):
    """MultiModalAttentionLayer.forward, src/model.py:181-202.
    topic [1,B,H], img/txt [S,B,H] -> [S,B,2048]."""
    S = img.shape[0]
    outs = []
    for i in range(S):
        w = sd[f"mm_atten_layer.att_matrices.{i}.weight"]
        b = sd[f"mm_atten_layer.att_matrices.{i}.bias"]
        cat = torch.stack([topic[0], img[i], txt[i]], 1)  # [B, 3, H]
        att = torch.softmax((cat @ w.t() + b).squeeze(-1), -1)  # [B, 3]
        o = (att.unsqueeze(-1) * cat).sum(1)  # [B, H]
        outs.append(o @ sd["mm_atten_layer.out_linear.weight"].t() + sd["mm_atten_layer.out_linear.bias"])
    return torch.stack(outs, 0)


def fused_context(sd, batch):
    """src/model.py:371-390: encoder -> 3 LayerNorms -> alpha x2 -> beta. Returns ([B,S,2048], kl)."""
    topic_emb = batch["topic_emb"].float()
    img = batch["img_embs"].transpose(0, 1).float()
    txt = batch["r_embs"].transpose(0, 1).float()
    t, i, x = encoder_forward(sd, topic_emb, img, txt)
    H = t.shape[-1]
    t = F.layer_norm(t, (H,), sd["ln_layer1.weight"], sd["ln_layer1.bias"], LN_EPS)
    i = F.layer_norm(i, (H,), sd["ln_layer2.weight"], sd["ln_layer2.bias"], LN_EPS)
    x = F.layer_norm(x, (H,), sd["ln_layer3.weight"], sd["ln_layer3.bias"], LN_EPS)
    ia, ikl = alpha_attention(sd, "img_inner_atten_layer", i.transpose(0, 1))
    xa, xkl = alpha_attention(sd, "text_inner_atten_layer", x.transpose(0, 1))
    mm = beta_attention(sd, t, ia.transpose(0, 1), xa.transpose(0, 1))
    return mm.transpose(0, 1), (ikl + xkl).mean()


# --------------------------------------------------------------------------------------------
# decoder
# --------------------------------------------------------------------------------------------
def gelu_new(x):
    """HF activations.py:59-66 NewGELUActivation."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def _drop(x, keep, p):
    """Inverted dropout with an INJECTED keep mask (torch.nn.functional.dropout's arithmetic)."""
    return x * keep.to(x.dtype) / (1.0 - p)


def gpt2_forward(sd, inputs_embeds, token_type_ids, attention_mask, n_layer=12, n_head=12,
                 prefix="decoder.gpt2.transformer.", dropout=None):
    """HF GPT2Model.forward + lm_head (modeling_gpt2.py:522-636, 703-706).
    Conv1D: y = x @ W + b with W stored [in, out] (pytorch_utils.py:97-123).
    dropout=None: eval mode. Otherwise the train-mode dropouts (embd :612, attention
    probabilities :67, resid after both c_proj :150/:259) with injected keep masks:
    {"p": (embd, resid, attn), "embd": [B,L,E], "attn": [n_layer x [B,NH,L,L]],
     "resid1": [n_layer x [B,L,E]], "resid2": [n_layer x [B,L,E]]}."""
    B, L, E = inputs_embeds.shape
    dh = E // n_head
    pos = torch.arange(L, device=inputs_embeds.device)
    h = inputs_embeds + sd[prefix + "wpe.weight"][pos]
    h = h + sd[prefix + "wte.weight"][token_type_ids]  # type ids index the WORD table
    if dropout is not None:
        pe, pr, pa = dropout["p"]
        h = _drop(h, dropout["embd"], pe)
    causal = torch.ones(L, L, dtype=torch.bool, device=h.device).tril()
    keep = causal.view(1, 1, L, L) & (attention_mask.view(B, 1, 1, L) != 0)
    for l in range(n_layer):
        p = f"{prefix}h.{l}."
        x = F.layer_norm(h, (E,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], LN_EPS)
        qkv = x @ sd[p + "attn.c_attn.weight"] + sd[p + "attn.c_attn.bias"]
        q, k, v = qkv.split(E, -1)
        q = q.view(B, L, n_head, dh).transpose(1, 2)
        k = k.view(B, L, n_head, dh).transpose(1, 2)
        v = v.view(B, L, n_head, dh).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
        s = s.masked_fill(~keep, float("-inf"))
        w = torch.softmax(s, -1)
        if dropout is not None:
            w = _drop(w, dropout["attn"][l], pa)
        a = w @ v
        a = a.transpose(1, 2).reshape(B, L, E)
        o = a @ sd[p + "attn.c_proj.weight"] + sd[p + "attn.c_proj.bias"]
        if dropout is not None:
            o = _drop(o, dropout["resid1"][l], pr)
        h = h + o
        x = F.layer_norm(h, (E,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], LN_EPS)
        u = gelu_new(x @ sd[p + "mlp.c_fc.weight"] + sd[p + "mlp.c_fc.bias"])
        o = u @ sd[p + "mlp.c_proj.weight"] + sd[p + "mlp.c_proj.bias"]
        if dropout is not None:
            o = _drop(o, dropout["resid2"][l], pr)
        h = h + o
    h = F.layer_norm(h, (E,), sd[prefix + "ln_f.weight"], sd[prefix + "ln_f.bias"], LN_EPS)
    return h @ sd[prefix + "wte.weight"].t()  # tied lm_head, no bias


def hf_causal_lm_loss(logits, labels):
    """HF ForCausalLMLoss (loss/loss_utils.py:28-67): mean CE of logits[:, :-1] vs labels[:, 1:];
    no label is -100 on this path, so PAD labels count."""
    V = logits.shape[-1]
    return F.cross_entropy(logits[:, :-1].reshape(-1, V).float(), labels[:, 1:].reshape(-1))


def decoder_embed(table, ctx, topic_ids, input_ids, two_sents_length):
    """src/model.py:253-268: token -> WenLan lookup, + fused context per sentence pair.
    table: [V, 2048] tensor; ctx [B, S, 2048]; returns ([B,P,2048], [B,T,2048])."""
    t_emb = table[topic_ids.long()]
    i_emb = table[input_ids.long()].clone()
    S = ctx.shape[1]
    T = input_ids.shape[1]
    for k in range(S):
        lo, hi = two_sents_length * k, min(two_sents_length * (k + 1), T)
        if lo < hi:
            i_emb[:, lo:hi] = i_emb[:, lo:hi] + ctx[:, k:k + 1]
    return t_emb, i_emb


def inference_type_ids_and_mask(input_ids_row0, tpw_type_ids, tpw_att_mask, sent_len, max_seq_length):
    """src/model.py:291-312 — derived from ROW 0 of input_ids only (reference behaviour)."""
    B = tpw_type_ids.shape[0]
    max_sent_num = max_seq_length // sent_len + 1
    tlist = list(range(1, max_sent_num)) + [1]
    types, mask = [], []
    for i, tok in enumerate(input_ids_row0.tolist()):
        if (i + 1) % sent_len == 0 or (i + 1) % sent_len == 1:
            types.append(0)
        else:
            types.append(0 if tok == 0 else tlist[i // sent_len])
        mask.append(0 if tok == 0 else 1)
    types = torch.tensor(types, dtype=torch.long).view(1, -1).expand(B, -1)
    mask = torch.tensor(mask, dtype=torch.long).view(1, -1).expand(B, -1)
    return (torch.cat([tpw_type_ids.long(), types.to(tpw_type_ids.device)], 1),
            torch.cat([tpw_att_mask.long(), mask.to(tpw_att_mask.device)], 1))


def mmtg_forward(sd, table, batch, data_config, train_flag=True, dropout=None):
    """MMTG.forward, src/model.py:356-400 -> (hf_loss, kl, logits [B, P+T, V]).
    `dropout`: injected GPT-2 keep masks (see gpt2_forward); None = eval mode."""
    ctx, kl = fused_context(sd, batch)
    input_ids, topic_ids = batch["targets"].long(), batch["topic_ids"].long()
    sent_len = data_config["max_sent_length"] + 2
    t_emb, i_emb = decoder_embed(table, ctx, topic_ids, input_ids, 2 * sent_len)
    emb = torch.cat([t_emb, i_emb], 1)
    if train_flag:
        types = torch.cat([batch["tpw_type_ids"].long(), batch["type_ids"].long()], 1)
        mask = torch.cat([batch["tpw_attention_mask"].long(), batch["attention_mask"].long()], 1)
        labels = torch.cat([topic_ids, input_ids], 1)
    else:
        types, mask = inference_type_ids_and_mask(input_ids[0], batch["tpw_type_ids"],
                                                  batch["tpw_attention_mask"], sent_len,
                                                  data_config["max_seq_length"])
        labels = torch.zeros(emb.shape[0], emb.shape[1], dtype=torch.long)
    h1 = torch.tanh(emb @ sd["decoder.projector_layer1.weight"].t() + sd["decoder.projector_layer1.bias"])
    x = h1 @ sd["decoder.projector_layer2.weight"].t() + sd["decoder.projector_layer2.bias"]
    n_layer = sum(1 for k in sd if k.startswith("decoder.gpt2.transformer.h.") and k.endswith(".ln_1.weight"))
    logits = gpt2_forward(sd, x, types, mask, n_layer=n_layer, dropout=dropout)
    return hf_causal_lm_loss(logits, labels), kl, logits


# --------------------------------------------------------------------------------------------
# loss
# --------------------------------------------------------------------------------------------
def my_loss(outputs, targets, ratings, stage, topic_prompt_length=15):
    """MyLoss.forward, src/loss.py:45-74."""
    NEAR_0 = 1e-10
    y = (ratings > 4) if stage == 1 else (ratings > 3)
    y = y.to(outputs.dtype)
    shift_logits = outputs[:, topic_prompt_length:-1, :]
    shift_labels = targets[:, 1:].long()
    losses = []
    for b in range(targets.shape[0]):
        ce = F.cross_entropy(shift_logits[b], shift_labels[b])
        p = 1 / torch.exp(ce)
        losses.append(-y[b] * torch.log(p + NEAR_0) - (1 - y[b]) * torch.log(1 - p + NEAR_0))
    return torch.stack(losses).mean()


# --------------------------------------------------------------------------------------------
# generation
# --------------------------------------------------------------------------------------------
def top_k_top_p_filtering(logits, top_k=0, top_p=0.0, filter_value=-float("inf")):
    """src/generate.py:64-94 (1-D logits, modified in place like the reference)."""
    assert logits.dim() == 1
    top_k = min(top_k, logits.size(-1))
    if top_k > 0:
        kth = torch.topk(logits, top_k)[0][-1]
        logits[logits < kth] = filter_value
    if top_p > 0.0:
        s_logits, s_idx = torch.sort(logits, descending=True)
        cum = torch.cumsum(torch.softmax(s_logits, -1), -1)
        remove = cum > top_p
        remove[1:] = remove[:-1].clone()
        remove[0] = False
        logits[s_idx[remove]] = filter_value
    return logits


def process_next_token_logits(logits, generated, temperature, repetition_penalty,
                              banned=(1, 2, 100, 102)):
    """src/generate.py:127-136. `set()` over 0-d tensors never de-duplicates, so the penalty is
    applied once per OCCURRENCE of a token id (ids 0 and 102 exempt), by plain division."""
    for tok in generated:
        if tok in (0, 102):
            continue
        logits[tok] = logits[tok] / repetition_penalty
    logits = logits / temperature
    for b in banned:
        logits[b] = -float("inf")
    return logits


def sample_sequence(sd, table, start_input, length, data_config, temperature=1.0, top_k=30,
                    top_p=0.0, repitition_penalty=1.0, generator=None, return_logits=False):
    """src/generate.py:97-145 for ONE sample (batch 1), full-prefix recompute per token.
    Returns the id list the reference returns (i.e. WITHOUT the last appended token)."""
    inputs = {}
    for k, v in start_input.items():
        if k == "targets":
            inputs[k] = torch.as_tensor(np.asarray(v), dtype=torch.long).unsqueeze(0)
        elif k != "rating":
            inputs[k] = torch.as_tensor(np.asarray(v), dtype=torch.float32).unsqueeze(0)
    sent_len = data_config["max_sent_length"] + 2
    generated = inputs["targets"]
    step_logits = []
    with torch.no_grad():
        for i in range(length):
            if i > 0 and (i + 2) % sent_len == 0:
                inputs["targets"] = torch.cat([inputs["targets"], torch.tensor([[2]])], -1)
                continue
            if i > 0 and (i + 2) % sent_len == 1:
                inputs["targets"] = torch.cat([inputs["targets"], torch.tensor([[1]])], -1)
                continue
            _, _, out = mmtg_forward(sd, table, inputs, data_config, train_flag=False)
            nxt = out[0, -1, :].clone()
            if return_logits:
                step_logits.append(nxt.clone())
            generated = inputs["targets"]
            nxt = process_next_token_logits(nxt, generated[0].tolist(), temperature, repitition_penalty)
            if generated[0, -1].item() == 0:
                tok = torch.tensor([[0]])
            else:
                filt = top_k_top_p_filtering(nxt, top_k=top_k, top_p=top_p)
                tok = torch.multinomial(torch.softmax(filt, -1), 1, generator=generator).unsqueeze(0)
            inputs["targets"] = torch.cat([generated, tok], -1)
    ids = generated.tolist()[0]
    return (ids, step_logits) if return_logits else ids
