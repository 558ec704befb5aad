"""ORACLE tooling — imports the UNMODIFIED reference: /root/reference/src in the build container,
else the byte-for-byte staged copy oracle/_ref/src (oracle/build_ref.py; git-ignored, travels to
the GPU box with the snapshot).

Used by scripts/make_golden.py (to write tests/golden/), tests/test_oracle_vs_reference.py
(skipped when neither tree is present) and bench.py's reference arm / cpu_baseline leg
(`kind: "reference"`): checker and baseline only, never on the product path.

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

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "src")
REF_SRC = "/root/reference/src" if os.path.isfile("/root/reference/src/model.py") else _STAGED


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
