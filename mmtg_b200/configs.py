"""Hyper-parameters of the MMTG hot path.

Values restate /root/reference/src/configs.py:14-54 and src/config/model_config.json:1-9 (they
are inputs of the path, not behaviour). `data_config` keeps the reference's dual attribute /
`__getitem__` access; unlike the reference it can be built for other sentence lengths
(BASELINE.json config 5: max_sent_length in {20, 40, 60, 98}).
"""
from __future__ import annotations

WENLAN_DIM = 2048
ENC_HIDDEN = 512


def _recurrent(kind: str = "GRU") -> dict:
    return {"type": kind, "input_dim": WENLAN_DIM, "hidden_dim": ENC_HIDDEN, "num_layers": 1}


def make_model_cfgs(seq_len: int = 5, heads: int = 4, dropout: float = 0.1) -> dict:
    return {
        "seq_len": seq_len,
        "topic": {"input_dim": WENLAN_DIM, "hidden_dim": ENC_HIDDEN},
        "image": _recurrent(),
        "text": _recurrent(),
        "SELF_ATT": {"hidden_size": ENC_HIDDEN, "attention_heads": heads},
        "MM_ATT": {"attention_dim": 1},
        "GPT2_PATH": "./pretrained/GPT2_lyrics_ckpt_epoch00.ckpt",
        "dropout": dropout,
    }


model_cfgs = make_model_cfgs()

GPT2_CONFIG = {
    "initializer_range": 0.02,
    "layer_norm_epsilon": 1e-05,
    "n_ctx": 250,
    "n_embd": 768,
    "n_head": 12,
    "n_layer": 12,
    "n_positions": 1024,
    "vocab_size": 13317,
    # HF GPT2Config defaults (configuration_gpt2.py); config/model_config.json does not override
    # them, so they are live in the reference's training forward
    "embd_pdrop": 0.1,
    "resid_pdrop": 0.1,
    "attn_pdrop": 0.1,
}


class data_config:
    """Sequence geometry: 10 sentences of (START + max_sent_length + EOS) tokens, + [SEP]."""

    def __init__(self, max_sent_length: int = 20, n_sentences: int = 10,
                 topic_prompt_length: int = 15, wenlan_emb_size: int = WENLAN_DIM):
        self.topic_prompt_length = topic_prompt_length
        self.max_sent_length = max_sent_length
        self.max_seq_length = n_sentences * (max_sent_length + 2)
        self.wenlan_emb_size = wenlan_emb_size

    def __getitem__(self, key):
        if not hasattr(self, key):
            print("No {} exists!".format(key))
            return None
        return getattr(self, key)
