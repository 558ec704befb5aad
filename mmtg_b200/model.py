"""Drop-in `MMTG` module (mirrors /root/reference/src/model.py:330-400) over the native engine.

Same constructor, `forward(batch) -> (loss, kl_loss, logits)` signature, `train_flag` attribute and
state_dict layout (193 entries) as the reference, but:
  * every parameter is a view into ONE flat fp32 buffer (plus a bf16 shadow for GEMM operands and a
    flat fp32 gradient buffer) so the engine addresses weights by offset, the gradient all-reduce
    works on contiguous buckets and the bf16 refresh is a single kernel;
  * forward/backward run in libmmtg_b200.so (hand-written sm_100a kernels); PyTorch only owns
    memory, streams and the autograd edge. There is no CPU / eager fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import pickle
import weakref

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .configs import GPT2_CONFIG

MAX_LAYERS = 48


# ----------------------------------------------------------------------------------------------
# ctypes mirrors of include/mmtg_b200.h
# ----------------------------------------------------------------------------------------------
class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("B", "P", "T", "L", "S", "two_sent", "Dw", "He", "alpha_heads", "E", "NH", "NL", "V",
                 "Vp", "n_pos", "_pad")]


class LayerOffsets(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ("ln1_w", "ln1_b", "attn_w", "attn_b", "proj_w", "proj_b", "ln2_w", "ln2_b", "fc_w",
                 "fc_b", "proj2_w", "proj2_b")]


class ParamOffsets(C.Structure):
    _fields_ = [
        ("topic_w", C.c_int64), ("topic_b", C.c_int64),
        ("gru_w_ih", C.c_int64 * 2), ("gru_w_hh", C.c_int64 * 2),
        ("gru_b_ih", C.c_int64 * 2), ("gru_b_hh", C.c_int64 * 2),
        ("enc_ln_w", C.c_int64 * 3), ("enc_ln_b", C.c_int64 * 3),
        ("alpha_qkv_w", C.c_int64 * 2), ("alpha_qkv_b", C.c_int64 * 2),
        ("beta_att_w", C.c_int64), ("beta_att_b", C.c_int64),
        ("beta_out_w", C.c_int64), ("beta_out_b", C.c_int64),
        ("proj1_w", C.c_int64), ("proj1_b", C.c_int64), ("proj2_w", C.c_int64), ("proj2_b", C.c_int64),
        ("wte", C.c_int64), ("wpe", C.c_int64), ("lnf_w", C.c_int64), ("lnf_b", C.c_int64),
        ("layer", LayerOffsets * MAX_LAYERS),
    ]


class Model(C.Structure):
    _fields_ = [("dims", Dims), ("off", ParamOffsets), ("params", C.c_void_p),
                ("params_bf16", C.c_void_p), ("grads", C.c_void_p), ("token_table", C.c_void_p),
                ("drop_seed", C.c_void_p), ("p_embd", C.c_float), ("p_resid", C.c_float),
                ("p_attn", C.c_float), ("table_rows", C.c_int32)]


class Batch(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("topic_ids", "targets", "type_ids", "attn_mask", "topic_emb", "img_embs", "txt_embs")]


# ----------------------------------------------------------------------------------------------
# parameter holders (names reproduce the reference's state_dict keys)
# ----------------------------------------------------------------------------------------------
class _Affine(nn.Module):
    """weight [out, in] (+ bias [out]) — stands in for nn.Linear / nn.LayerNorm / HF Conv1D."""

    def __init__(self, *wshape, bias=True, bias_len=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*wshape))
        if bias:
            self.bias = nn.Parameter(torch.empty(bias_len if bias_len is not None else wshape[0]))


class _GRU(nn.Module):
    def __init__(self, din, h):
        super().__init__()
        self.weight_ih_l0 = nn.Parameter(torch.empty(3 * h, din))
        self.weight_hh_l0 = nn.Parameter(torch.empty(3 * h, h))
        self.bias_ih_l0 = nn.Parameter(torch.empty(3 * h))
        self.bias_hh_l0 = nn.Parameter(torch.empty(3 * h))


class MultiModalEncoder(nn.Module):  # src/model.py:24-88
    def __init__(self, cfg):
        super().__init__()
        for m in ("image", "text"):
            if cfg[m]["type"] != "GRU" or cfg[m]["num_layers"] != 1:
                raise NotImplementedError("mmtg_b200 implements the reference configuration: 1-layer GRU encoders")
        h = cfg["topic"]["hidden_dim"]
        self.topic_fc = _Affine(h, cfg["topic"]["input_dim"])
        self.rnns_image = _GRU(cfg["image"]["input_dim"], h)
        self.rnns_text = _GRU(cfg["text"]["input_dim"], h)


class InnerModalAttentionLayer(nn.Module):  # src/model.py:91-161
    def __init__(self, cfg):
        super().__init__()
        h = cfg["SELF_ATT"]["hidden_size"]
        if h % cfg["SELF_ATT"]["attention_heads"] != 0:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)"
                             % (h, cfg["SELF_ATT"]["attention_heads"]))
        self.query, self.key, self.value = _Affine(h, h), _Affine(h, h), _Affine(h, h)


class MultiModalAttentionLayer(nn.Module):  # src/model.py:164-202
    def __init__(self, cfg):
        super().__init__()
        h = cfg["topic"]["hidden_dim"]
        if cfg["MM_ATT"]["attention_dim"] != 1:
            raise NotImplementedError("attention_dim must be 1 (reference configuration)")
        self.att_matrices = nn.ModuleList([_Affine(1, h) for _ in range(cfg["seq_len"])])
        self.out_linear = _Affine(2048, h)


class _Attn(nn.Module):
    def __init__(self, e):
        super().__init__()
        self.c_attn = _Affine(e, 3 * e, bias_len=3 * e)  # HF Conv1D: weight [in, out]
        self.c_proj = _Affine(e, e, bias_len=e)


class _MLP(nn.Module):
    def __init__(self, e):
        super().__init__()
        self.c_fc = _Affine(e, 4 * e, bias_len=4 * e)
        self.c_proj = _Affine(4 * e, e, bias_len=e)


class _Block(nn.Module):
    def __init__(self, e):
        super().__init__()
        self.ln_1 = _Affine(e)
        self.attn = _Attn(e)
        self.ln_2 = _Affine(e)
        self.mlp = _MLP(e)


class _Transformer(nn.Module):
    def __init__(self, g):
        super().__init__()
        e = g["n_embd"]
        self.wte = _Affine(g["vocab_size"], e, bias=False)
        self.wpe = _Affine(g["n_positions"], e, bias=False)
        self.h = nn.ModuleList([_Block(e) for _ in range(g["n_layer"])])
        self.ln_f = _Affine(e)


class _GPT2LMHead(nn.Module):
    def __init__(self, g):
        super().__init__()
        self.transformer = _Transformer(g)
        self.lm_head = nn.Module()
        self.lm_head.weight = self.transformer.wte.weight  # tied (HF tie_word_embeddings)


class GPT2_Decoder(nn.Module):  # src/model.py:205-327
    def __init__(self, data_config, model_name="uer/gpt2-chinese-cluecorpussmall",
                 config_path="config/model_config.json", gpt2_config=None, token_table=None,
                 token_table_path="./vocab/token_id2emb_dict.pkl"):
        super().__init__()
        self.data_config = data_config
        g = dict(GPT2_CONFIG)
        if gpt2_config is not None:
            g.update(gpt2_config)
        elif os.path.isfile(config_path):
            with open(config_path) as f:
                g.update(json.load(f))
        self.config = g
        self.projector_layer1 = _Affine(512, data_config["wenlan_emb_size"])
        self.projector_layer2 = _Affine(g["n_embd"], 512)
        self.gpt2 = _GPT2LMHead(g)
        self._table_host = None
        if token_table is not None:
            self.set_token_table(token_table)
        elif os.path.isfile(token_table_path):
            self.load_token_id2emb(token_table_path)

    def load_token_id2emb(self, path):
        with open(path, "rb") as f:
            self.set_token_table(pickle.load(f))
        return self._table_host

    def set_token_table(self, table):
        """dict {id -> vector} (the reference's pickle) or a dense [V, D] array."""
        if isinstance(table, dict):
            n = max(table.keys()) + 1
            dense = np.zeros((n, len(next(iter(table.values())))), np.float32)
            for k, v in table.items():
                dense[int(k)] = np.asarray(v, np.float32)
            table = dense
        self._table_host = torch.as_tensor(np.asarray(table, np.float32)).contiguous()
        self._table_dev = None


_ALIGN = 64  # elements; keeps every tensor 16-byte aligned in both the fp32 and bf16 buffers


class MMTG(nn.Module):
    def __init__(self, model_cfgs, data_config, vocab_size, train_flag=False, gpt2_config=None,
                 token_table=None):
        super().__init__()
        self.model_cfgs = model_cfgs
        self.data_config = data_config
        self.vocab_size = vocab_size
        self.encoder = MultiModalEncoder(model_cfgs)
        h = model_cfgs["topic"]["hidden_dim"]
        self.ln_layer1, self.ln_layer2, self.ln_layer3 = _Affine(h), _Affine(h), _Affine(h)
        self.img_inner_atten_layer = InnerModalAttentionLayer(model_cfgs)
        self.text_inner_atten_layer = InnerModalAttentionLayer(model_cfgs)
        self.mm_atten_layer = MultiModalAttentionLayer(model_cfgs)
        self.decoder = GPT2_Decoder(data_config, gpt2_config=gpt2_config, token_table=token_table)
        self.train_flag = train_flag
        self._flat = None       # (P, W16, G) flat buffers
        self._w16_fresh = False  # True only right after FusedAdamW.step() (see _refresh_bf16)
        self._ws = {}           # workspace cache keyed by dims tuple
        self._serial = 0
        self._anchor = None
        self.grad_sync = None   # optional parallel.GradSync
        g = self.decoder.config
        # GPT-2 dropout, live while `self.training and self.train_flag` (nn.Module.eval() or
        # set_dropout(0, 0, 0) turn it off: parity with the reference is defined at p = 0)
        self._drop_p = (float(g.get("embd_pdrop", 0.1)), float(g.get("resid_pdrop", 0.1)),
                        float(g.get("attn_pdrop", 0.1)))
        self._drop_seed = None  # device int64: masks are functions of (seed, site, element)
        self._drop_seed_init = 0x5EED
        self._build_layout()
        self.reset_parameters()
        if train_flag and os.path.isfile(model_cfgs.get("GPT2_PATH", "")):
            # src/model.py:345-354: warm-start the decoder from the pre-trained lyrics GPT-2
            sd = torch.load(model_cfgs["GPT2_PATH"], map_location="cpu")
            sd = sd.get("state_dict", sd)
            self.decoder.load_state_dict({k: v for k, v in sd.items() if not k.endswith((".attn.bias", ".attn.masked_bias"))},
                                         strict=False)

    # ------------------------------------------------------------------------------------------
    # layout
    # ------------------------------------------------------------------------------------------
    def _build_layout(self):
        g = self.decoder.config
        named = dict(self.named_parameters())
        order = []
        for l in range(g["n_layer"]):
            p = f"decoder.gpt2.transformer.h.{l}."
            order += [p + n for n in ("ln_1.weight", "ln_1.bias", "attn.c_attn.weight", "attn.c_attn.bias",
                                      "attn.c_proj.weight", "attn.c_proj.bias", "ln_2.weight", "ln_2.bias",
                                      "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias")]
        self._layers_end_name = order[-1]
        order += ["encoder.topic_fc.weight", "encoder.topic_fc.bias"]
        for m in ("image", "text"):
            order += [f"encoder.rnns_{m}.{n}" for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
        for i in (1, 2, 3):
            order += [f"ln_layer{i}.weight", f"ln_layer{i}.bias"]
        for m in ("img", "text"):  # q|k|v stacked contiguously -> one [3H, H] GEMM operand
            order += [f"{m}_inner_atten_layer.{n}.weight" for n in ("query", "key", "value")]
            order += [f"{m}_inner_atten_layer.{n}.bias" for n in ("query", "key", "value")]
        S = self.model_cfgs["seq_len"]
        order += [f"mm_atten_layer.att_matrices.{i}.weight" for i in range(S)]
        order += [f"mm_atten_layer.att_matrices.{i}.bias" for i in range(S)]
        order += ["mm_atten_layer.out_linear.weight", "mm_atten_layer.out_linear.bias",
                  "decoder.projector_layer1.weight", "decoder.projector_layer1.bias",
                  "decoder.projector_layer2.weight", "decoder.projector_layer2.bias",
                  "decoder.gpt2.transformer.wpe.weight", "decoder.gpt2.transformer.ln_f.weight",
                  "decoder.gpt2.transformer.ln_f.bias", "decoder.gpt2.transformer.wte.weight"]
        assert sorted(order) == sorted(named.keys()), "layout does not cover the parameter set"
        packed_after = set()  # names that must follow their predecessor without padding
        for m in ("img", "text"):
            packed_after |= {f"{m}_inner_atten_layer.{n}.{k}" for n in ("key", "value") for k in ("weight", "bias")}
        packed_after |= {f"mm_atten_layer.att_matrices.{i}.{k}" for i in range(1, S) for k in ("weight", "bias")}
        off, layout = 0, {}
        for name in order:
            if name not in packed_after:
                off = (off + _ALIGN - 1) // _ALIGN * _ALIGN
            layout[name] = (off, named[name].numel())
            off += named[name].numel()
        self._layout = layout
        self._order = order
        self._flat_numel = (off + _ALIGN - 1) // _ALIGN * _ALIGN
        end = layout[self._layers_end_name]
        self._layers_flat_end = end[0] + end[1]

    def layer_bucket(self, l):
        """[lo, hi) element range of GPT-2 block l in the flat buffers."""
        p = f"decoder.gpt2.transformer.h.{l}."
        lo = self._layout[p + "ln_1.weight"][0]
        last = self._layout[p + "mlp.c_proj.bias"]
        return lo, last[0] + last[1]

    def tail_bucket(self):
        return self._layers_flat_end, self._flat_numel

    def wte_range(self):
        """[lo, hi) element range of the tied wte / lm_head weight (last in the flat order) and its row width."""
        lo, n = self._layout["decoder.gpt2.transformer.wte.weight"]
        return lo, lo + n, self.decoder.config["n_embd"]

    def tail_buckets(self):
        """The tail split where the backward splits it: (projector, wpe, ln_f, tied wte) are final
        after stage NL+1, (encoder, multi-modal attention) after stage NL+2."""
        mid = self._layout["decoder.projector_layer1.weight"][0]
        return (mid, self._flat_numel), (self._layers_flat_end, mid)

    def reset_parameters(self):
        """Reference init: xavier/orthogonal encoder (src/model.py:83-88), nn.Linear defaults,
        HF GPT-2 normal(0, 0.02) (residual projections scaled by 1/sqrt(2 n_layer))."""
        g = self.decoder.config
        std = g["initializer_range"]
        named = dict(self.named_parameters())
        with torch.no_grad():
            for name, p in named.items():
                leaf = name.rsplit(".", 2)[-2:]
                if ".gpt2." in name:
                    if name.endswith("bias"):
                        p.zero_()
                    elif ".ln_" in name:
                        p.fill_(1.0)
                    elif name.endswith("c_proj.weight"):
                        p.normal_(0.0, std / math.sqrt(2 * g["n_layer"]))
                    else:
                        p.normal_(0.0, std)
                elif name.startswith("ln_layer"):
                    p.fill_(1.0) if name.endswith("weight") else p.zero_()
                elif name in ("encoder.topic_fc.weight",) or leaf[-1] == "weight_ih_l0":
                    nn.init.xavier_normal_(p)
                elif leaf[-1] == "weight_hh_l0":
                    nn.init.orthogonal_(p)
                elif leaf[-1] in ("bias_ih_l0", "bias_hh_l0"):
                    b = 1.0 / math.sqrt(p.numel() // 3)
                    p.uniform_(-b, b)
                elif name.endswith("weight"):
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                else:  # nn.Linear bias: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
                    w = named[name[:-4] + "weight"]
                    b = 1.0 / math.sqrt(w.shape[1])
                    p.uniform_(-b, b)

    # ------------------------------------------------------------------------------------------
    # flat buffers
    # ------------------------------------------------------------------------------------------
    def _ensure_flat(self, device):
        named = dict(self.named_parameters())
        ok = self._flat is not None and self._flat[0].device == device
        if ok:
            base = self._flat[0].data_ptr()
            first, last = self._order[0], self._order[-1]
            ok = (named[first].data_ptr() == base + 4 * self._layout[first][0]
                  and named[last].data_ptr() == base + 4 * self._layout[last][0])
        if ok:
            return
        if device.type != "cuda":
            raise _lib.MMTGError("mmtg_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        P = torch.zeros(self._flat_numel, device=device, dtype=torch.float32)
        G = torch.zeros(self._flat_numel, device=device, dtype=torch.float32)
        W16 = torch.zeros(self._flat_numel, device=device, dtype=torch.bfloat16)
        with torch.no_grad():
            for name, p in named.items():
                off, n = self._layout[name]
                view = P[off:off + n].view(p.shape)
                view.copy_(p.data.to(device))
                p.data = view
                p.grad = None
        self._flat = (P, W16, G)
        self._w16_fresh = False
        self._grads_fresh = True
        self._named = named
        self._offsets = self._make_offsets()

    def _make_offsets(self):
        o = ParamOffsets()
        L = self._layout
        at = lambda n: L[n][0]
        o.topic_w, o.topic_b = at("encoder.topic_fc.weight"), at("encoder.topic_fc.bias")
        for i, m in enumerate(("image", "text")):
            o.gru_w_ih[i] = at(f"encoder.rnns_{m}.weight_ih_l0")
            o.gru_w_hh[i] = at(f"encoder.rnns_{m}.weight_hh_l0")
            o.gru_b_ih[i] = at(f"encoder.rnns_{m}.bias_ih_l0")
            o.gru_b_hh[i] = at(f"encoder.rnns_{m}.bias_hh_l0")
        for i in range(3):
            o.enc_ln_w[i], o.enc_ln_b[i] = at(f"ln_layer{i + 1}.weight"), at(f"ln_layer{i + 1}.bias")
        for i, m in enumerate(("img", "text")):
            o.alpha_qkv_w[i] = at(f"{m}_inner_atten_layer.query.weight")
            o.alpha_qkv_b[i] = at(f"{m}_inner_atten_layer.query.bias")
        o.beta_att_w, o.beta_att_b = at("mm_atten_layer.att_matrices.0.weight"), at("mm_atten_layer.att_matrices.0.bias")
        o.beta_out_w, o.beta_out_b = at("mm_atten_layer.out_linear.weight"), at("mm_atten_layer.out_linear.bias")
        o.proj1_w, o.proj1_b = at("decoder.projector_layer1.weight"), at("decoder.projector_layer1.bias")
        o.proj2_w, o.proj2_b = at("decoder.projector_layer2.weight"), at("decoder.projector_layer2.bias")
        t = "decoder.gpt2.transformer."
        o.wte, o.wpe = at(t + "wte.weight"), at(t + "wpe.weight")
        o.lnf_w, o.lnf_b = at(t + "ln_f.weight"), at(t + "ln_f.bias")
        for l in range(self.decoder.config["n_layer"]):
            p, lo = f"{t}h.{l}.", o.layer[l]
            lo.ln1_w, lo.ln1_b = at(p + "ln_1.weight"), at(p + "ln_1.bias")
            lo.attn_w, lo.attn_b = at(p + "attn.c_attn.weight"), at(p + "attn.c_attn.bias")
            lo.proj_w, lo.proj_b = at(p + "attn.c_proj.weight"), at(p + "attn.c_proj.bias")
            lo.ln2_w, lo.ln2_b = at(p + "ln_2.weight"), at(p + "ln_2.bias")
            lo.fc_w, lo.fc_b = at(p + "mlp.c_fc.weight"), at(p + "mlp.c_fc.bias")
            lo.proj2_w, lo.proj2_b = at(p + "mlp.c_proj.weight"), at(p + "mlp.c_proj.bias")
        return o

    def _refresh_bf16(self):
        """Re-cast the bf16 weight shadow from the fp32 masters at EVERY forward, unless the one
        writer that keeps the shadow coherent itself (FusedAdamW: its kernel stores the bf16 copy
        next to the fp32 update) has marked it fresh since the previous forward. Tensor version
        counters are NOT trusted: `p.data.add_()` (transformers.AdamW of the pinned 4.12.3, EMA,
        clamping, manual copies) changes the masters without bumping them. One HBM-bound kernel
        (436 MB read + 218 MB written, ~0.1 ms)."""
        if self._w16_fresh:
            self._w16_fresh = False  # consumed: whatever happens before the next forward is unknown
            return
        P, W16, _ = self._flat
        _lib.check(_lib.lib().mmtg_cast_bf16(C.c_void_p(P.data_ptr()), C.c_void_p(W16.data_ptr()),
                                             C.c_int64(P.numel()), C.c_void_p(_lib.stream_ptr())),
                   "mmtg_cast_bf16")

    def mark_bf16_shadow_fresh(self):
        """Called by FusedAdamW.step(): the optimizer kernel has just rewritten the bf16 shadow."""
        self._w16_fresh = True

    def _table(self, device):
        d = self.decoder
        if d._table_host is None:
            raise _lib.MMTGError("token table missing: pass token_table=... or call "
                                 "model.decoder.set_token_table(...) / load_token_id2emb(path)")
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:  # "cuda" and "cuda:<current>" are one device
            device = torch.device("cuda", torch.cuda.current_device())
        if getattr(d, "_table_dev", None) is None or d._table_dev.device != device:
            d._table_dev = d._table_host.to(device)
        return d._table_dev

    def _dims(self, B, T):
        dc, g, cfg = self.data_config, self.decoder.config, self.model_cfgs
        d = Dims()
        d.B, d.P, d.T = B, dc["topic_prompt_length"], T
        d.L = d.P + d.T
        d.S, d.two_sent = cfg["seq_len"], 2 * (dc["max_sent_length"] + 2)
        d.Dw, d.He, d.alpha_heads = dc["wenlan_emb_size"], cfg["topic"]["hidden_dim"], cfg["SELF_ATT"]["attention_heads"]
        d.E, d.NH, d.NL, d.V = g["n_embd"], g["n_head"], g["n_layer"], g["vocab_size"]
        d.Vp = (d.V + 7) // 8 * 8
        d.n_pos = g["n_positions"]
        return d

    def _workspace(self, d, device):
        key = (d.B, d.T, device)
        ws = self._ws.get(key)
        if ws is None:
            n = _lib.lib().mmtg_train_workspace_bytes(C.byref(d))
            if n < 0:
                _lib.check(-1, "mmtg_train_workspace_bytes")
            if len(self._ws) > 4:
                self._ws.clear()
            ws = torch.empty(n, dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    def forward(self, batch):
        """batch: the reference's dict (src/model.py:357-370). Returns (hf_loss, kl_loss, logits)."""
        img = batch["img_embs"]
        device = img.device
        self._ensure_flat(device)
        self._refresh_bf16()
        B, T = img.size(0), batch["targets"].size(1)
        d = self._dims(B, T)
        P = d.P
        targets = batch["targets"].long()
        topic_ids = batch["topic_ids"].long()
        if self.train_flag:
            types = torch.cat([batch["tpw_type_ids"].long(), batch["type_ids"].long()], 1)
            mask = torch.cat([batch["tpw_attention_mask"].long(), batch["attention_mask"].long()], 1)
        else:
            types, mask = _inference_types_and_mask(targets, batch["tpw_type_ids"], batch["tpw_attention_mask"],
                                                    self.data_config)
        floats = [batch["topic_emb"].float().contiguous(), img.float().contiguous(),
                  batch["r_embs"].float().contiguous()]
        step = _Step()
        step.model, step.dims, step.B, step.T = self, d, B, T
        step.floats = floats
        step.topic_ids = topic_ids.to(torch.int32).contiguous()
        step.targets = targets.to(torch.int32).contiguous()
        step.types = types.to(torch.int32).contiguous()
        step.mask = mask.to(torch.int32).contiguous()
        assert P == topic_ids.size(1), "topic_ids width must equal data_config.topic_prompt_length"
        step.ws = self._workspace(d, device)
        self._serial += 1
        step.serial = self._serial
        step.ws_serial_owner = self
        step.scalars = torch.zeros(4, device=device, dtype=torch.float32)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._named.values())
        if need_grad:
            if self._anchor is None or self._anchor.device != device:
                self._anchor = torch.zeros((), device=device, requires_grad=True)
            hf, kl, logits, token = _MMTGFunction.apply(self._anchor, step)
            step.token = token
            step.logits = logits  # weak (see _Step)
            logits._mmtg_step = step
            return hf, kl, logits
        logits = torch.empty(B, d.L, d.V, device=device, dtype=torch.float32)
        _run_forward(step, logits)
        step.logits = logits
        logits._mmtg_step = step
        step.token = None
        return step.scalars[0], step.scalars[1], logits

    # ------------------------------------------------------------------------------------------
    def fused_forward_loss(self, batch, stage, alpha=0.2, grad_scale=1.0):
        """forward + MyLoss + the bf16 dlogits operand, WITHOUT torch.autograd: the same engine
        entry points `forward()` / `MyLoss` use, driven directly (restates src/train.py:188-192).
        Returns (step, total, loss, kl); `backward_stages(step, ...)` completes the step.
        `grad_scale` multiplies d(total) (ragged data-parallel batches: B_local / B_global with a
        SUM all-reduce)."""
        prev = torch.is_grad_enabled()
        torch.set_grad_enabled(False)
        try:
            _, _, logits = self.forward(batch)
        finally:
            torch.set_grad_enabled(prev)
        step = logits._mmtg_step
        d, dev = step.dims, logits.device
        lib, st = _lib.lib(), C.c_void_p(_lib.stream_ptr())
        vp = C.c_void_p
        ratings = batch["rating"].to(device=dev, dtype=torch.int32).contiguous()
        ce, coef = torch.empty(d.B, device=dev), torch.empty(d.B, device=dev)
        loss = torch.empty((), device=dev)
        _lib.check(lib.mmtg_ce_reduce(vp(logits.data_ptr()), C.c_int64(d.V), vp(step.lse_ptr), None,
                                      vp(step.targets.data_ptr()), None, vp(ce.data_ptr()), None, d.B, d.L,
                                      d.P, d.T, st), "mmtg_ce_reduce")
        _lib.check(lib.mmtg_negloss(vp(ce.data_ptr()), vp(ratings.data_ptr()), int(stage), vp(loss.data_ptr()),
                                    vp(coef.data_ptr()), d.B, st), "mmtg_negloss")
        kl = step.scalars[1]
        total = loss + alpha * kl
        if getattr(self, "_unit", None) is None or self._unit.device != dev:
            self._unit = torch.ones(1, device=dev)
            self._alpha_buf = torch.zeros(1, device=dev)
            self._alpha_val = None
        if self._alpha_val != (alpha, grad_scale):
            self._unit.fill_(grad_scale)
            self._alpha_buf.fill_(alpha * grad_scale)
            self._alpha_val = (alpha, grad_scale)
        _lib.check(lib.mmtg_ce_bwd(vp(logits.data_ptr()), C.c_int64(d.V), vp(step.lse_ptr), None,
                                   vp(step.targets.data_ptr()), vp(coef.data_ptr()), vp(self._unit.data_ptr()),
                                   None, vp(step.dlogits_ptr), 1, C.c_int64(d.Vp), d.B, d.L, d.P, d.T, d.V, st),
                   "mmtg_ce_bwd")
        step.keep = (ratings, ce, coef)
        _attach_grads(self)
        return step, total, loss, kl

    def backward_stages(self, step, s0=0, s1=None):
        """Run backward stages [s0, s1) of `step` (0 = lm_head, 1..NL = blocks NL-1..0, NL+1 = embeddings/projector, NL+2 =
        embeddings/encoder) with d(total)/d(kl) = alpha; gradients accumulate into param.grad."""
        d = step.dims
        s1 = d.NL + 3 if s1 is None else s1
        _lib.check(_lib.lib().mmtg_train_backward(C.byref(step.cm), C.byref(step.cb), C.c_void_p(step.ws.data_ptr()),
                                                  C.c_int64(step.ws.numel()), C.c_void_p(self._alpha_buf.data_ptr()),
                                                  s0, s1, C.c_void_p(_lib.stream_ptr())), "mmtg_train_backward")

    def fused_train_step(self, batch, stage, alpha=0.2, grad_scale=1.0):
        """fused_forward_loss + all backward stages (+ bucketed gradient all-reduce when
        `grad_sync` is set). No Python-side synchronisation and no autograd graph, so the call is
        CUDA-graph capturable (mmtg_b200.graph.GraphedTrainStep). Returns (total, loss, kl)."""
        step, total, loss, kl = self.fused_forward_loss(batch, stage, alpha, grad_scale)
        _run_backward(step, self._alpha_buf)
        return total, loss, kl

    def set_dropout(self, embd=None, resid=None, attn=None):
        """Override the GPT-2 dropout probabilities (HF embd_pdrop / resid_pdrop / attn_pdrop)."""
        e, r, a = self._drop_p
        self._drop_p = (e if embd is None else float(embd), r if resid is None else float(resid),
                        a if attn is None else float(attn))
        return self

    def set_dropout_seed(self, seed):
        """Seed of the counter-based dropout masks; every training forward advances it."""
        self._drop_seed_init = int(seed) & 0x7FFFFFFFFFFFFFFF
        if self._drop_seed is not None:
            self._drop_seed.fill_(self._drop_seed_init)
        return self

    def dropout_active(self):
        return bool(self.training and self.train_flag and any(p > 0 for p in self._drop_p))

    def _seed_tensor(self, device):
        if self._drop_seed is None or self._drop_seed.device != device:
            seed = self._drop_seed_init
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                seed += 7919 * torch.distributed.get_rank()  # independent masks per data-parallel rank
            self._drop_seed = torch.full((1,), seed, dtype=torch.int64, device=device)
        return self._drop_seed

    def _c_model(self, d, device, train=False):
        m = Model()
        m.dims, m.off = d, self._offsets
        P, W16, G = self._flat
        m.params, m.params_bf16, m.grads = P.data_ptr(), W16.data_ptr(), G.data_ptr()
        tab = self._table(device)
        m.token_table, m.table_rows = tab.data_ptr(), tab.shape[0]
        if train and self.dropout_active():
            m.drop_seed = self._seed_tensor(device).data_ptr()
            m.p_embd, m.p_resid, m.p_attn = self._drop_p
        return m

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts the reference's checkpoints: optional DataParallel `module.` prefix
        (src/train.py:113,212) and transformers-4.x causal-mask buffers (`attn.bias`,
        `attn.masked_bias`) are tolerated (SURVEY §8b caveats)."""
        sd = {}
        for k, v in state_dict.items():
            if k.startswith("module."):
                k = k[len("module."):]
            if k.endswith((".attn.bias", ".attn.masked_bias")):
                continue
            sd[k] = v
        self._w16_fresh = False  # the masters change underneath the bf16 shadow
        return super().load_state_dict(sd, strict=strict, **kw)


def _inference_types_and_mask(input_ids, tpw_type_ids, tpw_att_mask, data_config):
    """Inference-branch token types / key mask (src/model.py:291-312): derived from ROW 0 only,
    per-sentence type ids [1..10, 1], 0 at sentence boundaries and PAD."""
    sent_len = data_config["max_sent_length"] + 2
    max_sent_num = data_config["max_seq_length"] // sent_len + 1
    tlist = torch.tensor(list(range(1, max_sent_num)) + [1], device=input_ids.device)
    T = input_ids.size(1)
    i = torch.arange(T, device=input_ids.device)
    row0 = input_ids[0]
    boundary = ((i + 1) % sent_len == 0) | ((i + 1) % sent_len == 1)
    ty = torch.where(boundary | (row0 == 0), torch.zeros_like(i), tlist[(i // sent_len).clamp(max=tlist.numel() - 1)])
    mk = (row0 != 0).long()
    Bn = tpw_type_ids.size(0)
    return (torch.cat([tpw_type_ids.long(), ty.view(1, -1).expand(Bn, -1)], 1),
            torch.cat([tpw_att_mask.long(), mk.view(1, -1).expand(Bn, -1)], 1))


class _Step:
    """Everything one forward produced that backward / the fused loss needs.

    Ownership is one-directional: the logits tensor owns its step (`logits._mmtg_step`), the step
    only holds a WEAK reference back (`step.logits`), and the autograd nodes keep the tensors they
    need through `save_for_backward`. There is no reference cycle, so the 402 MB logits buffer of
    a B = 32 step is released by reference counting as soon as the caller drops it."""
    dlogits_ready = False
    _logits_ref = None

    @property
    def logits(self):
        return self._logits_ref() if self._logits_ref is not None else None

    @logits.setter
    def logits(self, t):
        self._logits_ref = weakref.ref(t) if t is not None else None


def _c_batch(step):
    b = Batch()
    b.topic_ids, b.targets = step.topic_ids.data_ptr(), step.targets.data_ptr()
    b.type_ids, b.attn_mask = step.types.data_ptr(), step.mask.data_ptr()
    b.topic_emb, b.img_embs, b.txt_embs = (t.data_ptr() for t in step.floats)
    return b


def _run_forward(step, logits):
    mdl = step.model
    device = logits.device
    m, b = mdl._c_model(step.dims, device, train=True), _c_batch(step)
    step.cm, step.cb = m, b
    if m.drop_seed:  # fresh masks for this step; backward regenerates them from the same seed
        _lib.check(_lib.lib().mmtg_dropout_next_seed(C.c_void_p(m.drop_seed), C.c_void_p(_lib.stream_ptr())),
                   "mmtg_dropout_next_seed")
    rc = _lib.lib().mmtg_train_forward(C.byref(m), C.byref(b), C.c_void_p(step.ws.data_ptr()),
                                       C.c_int64(step.ws.numel()), C.c_void_p(logits.data_ptr()),
                                       C.c_void_p(step.scalars.data_ptr()), 1, C.c_void_p(_lib.stream_ptr()))
    _lib.check(rc, "mmtg_train_forward")
    lib = _lib.lib()
    lib.mmtg_ws_lse.restype = C.c_void_p
    lib.mmtg_ws_dlogits_bf16.restype = C.c_void_p
    step.lse_ptr = lib.mmtg_ws_lse(C.byref(step.dims), C.c_void_p(step.ws.data_ptr()))
    step.dlogits_ptr = lib.mmtg_ws_dlogits_bf16(C.byref(step.dims), C.c_void_p(step.ws.data_ptr()))


class _MMTGFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, step):
        ctx.set_materialize_grads(False)
        ctx.step = step  # step -> logits is a weak reference: no cycle through the node
        d = step.dims
        logits = torch.empty(d.B, d.L, d.V, device=anchor.device, dtype=torch.float32)
        _run_forward(step, logits)
        token = torch.zeros((), device=logits.device)
        ctx.save_for_backward(logits)  # an output: autograd stores it without a cycle
        return step.scalars[0].clone(), step.scalars[1].clone(), logits, token

    @staticmethod
    def backward(ctx, g_hf, g_kl, g_logits, g_token):
        step = ctx.step
        mdl = step.model
        if step.serial != mdl._serial:
            raise _lib.MMTGError("backward() after a newer forward(): the activation workspace was overwritten")
        lib = _lib.lib()
        d = step.dims
        st = C.c_void_p(_lib.stream_ptr())
        wsp = C.c_void_p(step.ws.data_ptr())
        dense = g_logits
        if g_hf is not None:
            # HF loss gradient (train.py discards this loss; supported for completeness)
            (logits,) = ctx.saved_tensors
            hf_dense = torch.empty_like(logits)
            g = g_hf.detach().float().contiguous()
            _lib.check(lib.mmtg_ce_bwd(C.c_void_p(logits.data_ptr()), C.c_int64(d.V), C.c_void_p(step.lse_ptr),
                                       C.c_void_p(step.topic_ids.data_ptr()), C.c_void_p(step.targets.data_ptr()),
                                       None, None, C.c_void_p(g.data_ptr()), C.c_void_p(hf_dense.data_ptr()), 0,
                                       C.c_int64(d.V), d.B, d.L, d.P, d.T, d.V, st), "mmtg_ce_bwd")
            dense = hf_dense if dense is None else dense + hf_dense
        if dense is not None:
            if step.dlogits_ready:
                raise _lib.MMTGError("logits received both a fused-loss gradient and a dense gradient")
            dense = dense.detach().float().contiguous()
            _lib.check(lib.mmtg_dlogits_from_f32(C.byref(d), wsp, C.c_void_p(dense.data_ptr()), st),
                       "mmtg_dlogits_from_f32")
            step.dlogits_ready = True
        if not step.dlogits_ready:
            _zero_dlogits(step)  # only the KL term reaches the model
        gkl = None
        if g_kl is not None:
            gkl = g_kl.detach().float().contiguous()
        _attach_grads(mdl)
        _run_backward(step, gkl)
        step.dlogits_ready = False
        return None, None


def _attach_grads(mdl):
    """Gradients are written straight into the flat gradient buffer: param.grad = views of it."""
    P, W16, G = mdl._flat
    if any(p.grad is None for p in mdl._named.values()):
        G.zero_()
        for name, p in mdl._named.items():
            off, n = mdl._layout[name]
            p.grad = G[off:off + n].view(p.shape)


def _run_backward(step, gkl):
    """All backward stages; after each stage the finished gradient bucket is handed to the
    (optional) data-parallel GradSync, which all-reduces it on a side stream."""
    mdl, d = step.model, step.dims
    lib = _lib.lib()
    st = C.c_void_p(_lib.stream_ptr())
    wsp = C.c_void_p(step.ws.data_ptr())
    nstage = d.NL + 3
    sync = mdl.grad_sync
    if sync is None:
        _lib.check(lib.mmtg_train_backward(C.byref(step.cm), C.byref(step.cb), wsp, C.c_int64(step.ws.numel()),
                                           C.c_void_p(gkl.data_ptr()) if gkl is not None else None, 0, nstage, st),
                   "mmtg_train_backward")
        return
    for s in range(nstage):
        sync.before_stages(mdl, s, s + 1, nstage)
        rc = lib.mmtg_train_backward(C.byref(step.cm), C.byref(step.cb), wsp, C.c_int64(step.ws.numel()),
                                     C.c_void_p(gkl.data_ptr()) if gkl is not None else None, s, s + 1, st)
        _lib.check(rc, "mmtg_train_backward")
        sync.after_stage(mdl, s, nstage)
    sync.finish(mdl)


def _zero_dlogits(step):
    d = step.dims
    n = d.B * d.L * d.Vp
    base = step.ws.data_ptr()
    off = step.dlogits_ptr - base
    step.ws[off:off + 2 * n].zero_()
    step.dlogits_ready = True
