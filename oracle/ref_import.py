"""ORACLE tooling — imports the UNMODIFIED reference (/root/reference/src) in the build container.

/root/reference does not exist on the GPU box, so nothing under tests -m gpu, smoke() or
bench.py may call this; it is used by scripts/make_golden.py (to write tests/golden/) and by
tests/test_oracle_vs_reference.py (skipped when the reference tree is absent).

Recipe: a scratch working directory holding config/model_config.json and a
synthetic vocab/token_id2emb_dict.pkl (the constructor hard-codes both relative paths,
src/model.py:210,215), and GPT2LMHeadModel.from_pretrained stubbed to a random-init model of
that config (no network).
"""
from __future__ import annotations

import contextlib
import os
import pickle
import shutil
import sys
import tempfile

REF_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "model.py"))


@contextlib.contextmanager
def _chdir(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def load_reference(token_table, state_dict=None):
    """Returns (model, MyLoss instance, generate module, data_config instance, model_cfgs)."""
    import torch
    import transformers
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    work = tempfile.mkdtemp(prefix="mmtg_ref_")
    os.makedirs(os.path.join(work, "config"))
    os.makedirs(os.path.join(work, "vocab"))
    shutil.copy(os.path.join(REF_SRC, "config", "model_config.json"), os.path.join(work, "config"))
    with open(os.path.join(work, "vocab", "token_id2emb_dict.pkl"), "wb") as f:
        pickle.dump({i: token_table[i] for i in range(token_table.shape[0])}, f)

    def _from_pretrained(*_a, **_k):
        cfg = transformers.GPT2Config.from_json_file("config/model_config.json")
        return transformers.GPT2LMHeadModel(cfg)

    orig = transformers.GPT2LMHeadModel.from_pretrained
    transformers.GPT2LMHeadModel.from_pretrained = staticmethod(_from_pretrained)
    try:
        with _chdir(work):
            import configs as ref_configs
            import generate as ref_generate
            import loss as ref_loss
            import model as ref_model
            dc = ref_configs.data_config()
            m = ref_model.MMTG(ref_configs.model_cfgs, dc, 13317, train_flag=False)
    finally:
        transformers.GPT2LMHeadModel.from_pretrained = orig
        shutil.rmtree(work, ignore_errors=True)
    m.train_flag = True
    if state_dict is not None:
        missing, unexpected = m.load_state_dict(state_dict, strict=True)
        assert not missing and not unexpected
    m.eval()
    crit = ref_loss.MyLoss(dc, ref_configs.model_cfgs)
    return m, crit, ref_generate, dc, ref_configs.model_cfgs
