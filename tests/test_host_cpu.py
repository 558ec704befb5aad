"""CPU-side checks: the C-ABI library loads and exports every symbol include/mmtg_b200.h declares
(no compute without a GPU), the drop-in module reproduces the reference's state_dict layout, and
host-side logic (inference-branch type ids / masks, returned-length rule, synthetic batch layout)
matches the oracle / reference rules."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mmtg_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mmtg_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(mmtg_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    lib = _lib.lib()
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert lib.mmtg_abi_version() == 3
    lib.mmtg_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.mmtg_last_error(), bytes)


def test_ctypes_structs_match_header_sizes(tmp_path):
    """The ctypes mirrors against the C compiler's view of include/mmtg_b200.h (sizes and the
    offsets of the last fields), not against hand-counted constants."""
    import subprocess
    from mmtg_b200 import _lib
    from mmtg_b200.model import Batch, Dims, LayerOffsets, Model, ParamOffsets
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "mmtg_b200.h"\n'
        "int main(void) {\n"
        '  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(mmtg_dims), sizeof(mmtg_layer_offsets),\n'
        "         sizeof(mmtg_param_offsets), sizeof(mmtg_batch), sizeof(mmtg_model), sizeof(mmtg_gemm_args),\n"
        "         offsetof(mmtg_model, p_attn), offsetof(mmtg_gemm_args, drop_p), offsetof(mmtg_gemm_args, grid_mode));\n"
        "  return 0;\n}\n")
    exe = tmp_path / "sz"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-I", inc, str(src), "-o", str(exe)], check=True)
    c = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert ctypes.sizeof(Dims) == c[0] == 16 * 4
    assert ctypes.sizeof(LayerOffsets) == c[1]
    assert ctypes.sizeof(ParamOffsets) == c[2]
    assert ctypes.sizeof(Batch) == c[3]
    assert ctypes.sizeof(Model) == c[4]
    assert ctypes.sizeof(_lib.GemmArgs) == c[5]
    assert Model.p_attn.offset == c[6]
    assert _lib.GemmArgs.drop_p.offset == c[7] and _lib.GemmArgs.grid_mode.offset == c[8]


def test_workspace_size_query_runs_without_gpu():
    from mmtg_b200 import _lib
    from mmtg_b200.configs import data_config, model_cfgs
    from mmtg_b200.model import MMTG
    m = MMTG(model_cfgs, data_config(), 13317)
    d = m._dims(32, 221)
    n = _lib.lib().mmtg_train_workspace_bytes(ctypes.byref(d))
    assert 2e9 < n < 8e9, n
    d.E = 700  # head_dim != 64 -> rejected with a message, not a crash
    assert _lib.lib().mmtg_train_workspace_bytes(ctypes.byref(d)) == -1
    assert b"head_dim" in _lib.lib().mmtg_last_error()


def test_state_dict_layout_and_flat_layout():
    from mmtg_b200 import synth
    from mmtg_b200.configs import data_config, model_cfgs
    from mmtg_b200.model import MMTG
    m = MMTG(model_cfgs, data_config(), 13317)
    sd = m.state_dict()
    assert list(sd.keys()) == synth.state_dict_keys() and len(sd) == 193
    assert sum(p.numel() for p in m.parameters()) == 109_064_709
    assert sd["decoder.gpt2.lm_head.weight"].data_ptr() == sd["decoder.gpt2.transformer.wte.weight"].data_ptr()
    ref = synth.make_state_dict(0)
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    # q|k|v stacked contiguously; every tensor 64-element aligned; blocks form contiguous buckets
    lay = m._layout
    for pre in ("img", "text"):
        q, k, v = (lay[f"{pre}_inner_atten_layer.{n}.weight"] for n in ("query", "key", "value"))
        assert k[0] == q[0] + q[1] and v[0] == k[0] + k[1]
    spans = sorted(lay.values())
    for (o1, n1), (o2, _) in zip(spans, spans[1:]):
        assert o1 + n1 <= o2
    lo, hi = m.layer_bucket(3)
    assert hi - lo >= 7_087_872 and m.layer_bucket(4)[0] >= hi
    assert m.tail_bucket()[1] == m._flat_numel
    # no CPU path: forward on CPU tensors fails loudly
    batch = synth.batch_to_torch(synth.make_batch(1))
    with pytest.raises(Exception):
        m(batch)


def test_inference_types_and_mask_match_oracle():
    from mmtg_b200.configs import data_config
    from mmtg_b200.model import _inference_types_and_mask
    from oracle import mmtg_oracle as O
    dc = data_config()
    rng = np.random.default_rng(0)
    for T in (1, 5, 22, 23, 45, 130, 221):
        ids = torch.from_numpy(rng.integers(0, 300, (2, T)))
        ids[0, rng.integers(0, T)] = 0
        tt = torch.from_numpy(rng.integers(0, 2, (2, 15)))
        tm = torch.from_numpy(rng.integers(0, 2, (2, 15)))
        a = _inference_types_and_mask(ids, tt, tm, dc)
        b = O.inference_type_ids_and_mask(ids[0], tt, tm, 22, 220)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), T


def test_returned_length_rule():
    from mmtg_b200.generate import returned_length
    assert returned_length(220, 22) == 218   # reference: i = 218, 219 are forced slots
    assert returned_length(50, 22) == 50     # [probed] in SURVEY §3.4: length=50 -> 50 ids
    assert returned_length(48, 22) == 48
    assert returned_length(21, 22) == 20
    assert returned_length(1, 22) == 1


def test_synthetic_batch_follows_dataset_layout():
    from mmtg_b200 import synth
    b = synth.make_batch(4, seed=3)
    assert b["targets"].shape == (4, 221) and b["topic_ids"].shape == (4, 15)
    t, m, ty = b["targets"], b["attention_mask"], b["type_ids"]
    assert (t[:, 220] == 102).all() and (m[:, 220] == 1).all()
    for s in range(10):
        assert (t[:, 22 * s] == 1).all() and (t[:, 22 * s + 21] == 2).all()
        body = slice(22 * s + 1, 22 * s + 21)
        assert ((t[:, body] == 0) == (m[:, body] == 0)).all()
        want = 1 if s // 2 == 4 else s // 2 + 1
        assert set(np.unique(ty[:, body])) <= {0, want}
    assert (ty[:, ::22][:, :10] == 0).all()
    assert np.allclose(np.linalg.norm(b["img_embs"], axis=-1), 1.0, atol=1e-5)
    assert b["rating"].min() >= 1 and b["rating"].max() <= 5


def test_curriculum_filtering_matches_reference_rules():
    from mmtg_b200.curriculum import filter_batch, stage_for_epoch, stage_row_indices
    r = torch.tensor([3, 1, 5, 2, 4, 3, 5, 1])
    assert stage_row_indices(r, 1).tolist() == [1, 7, 2, 6]        # r<2 first, then r>4
    assert stage_row_indices(r, 2).tolist() == [1, 3, 7, 2, 4, 6]  # r<3 first, then r>3
    assert stage_row_indices(r, 3).tolist() == list(range(8))
    assert [stage_for_epoch(e, [1, 3]) for e in range(5)] == [1, 2, 2, 3, 3]  # train.sh: curriculums [1,3]
    batch = {"rating": r, "targets": torch.arange(16).view(8, 2)}
    out = filter_batch(batch, 1)
    assert out["rating"].tolist() == [1, 1, 5, 5] and out["targets"][0].tolist() == [2, 3]
    assert filter_batch({"rating": torch.tensor([3, 3])}, 2) is None


def test_postprocess_tokens():
    from mmtg_b200.generate import postprocess_tokens
    sent = ["[#START#]", "a", "b", "[PAD]", "[#EOS#]"]
    assert postprocess_tokens(sent * 10 + ["x", "y"]) == "，".join(["ab"] * 10)
    assert postprocess_tokens(sent * 2 + ["[SEP]", "z"]) == "ab，ab"
    assert postprocess_tokens(sent) == "ab"


def test_generate_samples_front_end_packs_and_orders_rows(tmp_path):
    """src/generate.py:203-244 restated over the batched decoder: n_samples lines per item, dataset
    order, every run started from [#START#], rows packed rows_per_call at a time."""
    import numpy as np
    from mmtg_b200.generate import generate_samples

    vocab = ["[PAD]", "[#START#]", "[#EOS#]", "a", "b", "[SEP]"]

    class Tok:
        def convert_tokens_to_ids(self, t):
            return vocab.index(t)

        def convert_ids_to_tokens(self, ids):
            return [vocab[i] for i in ids]

    class Model:
        data_config = {"max_seq_length": 7}

    calls = []

    def fake_sampler(model, starts, length, tokenizer, temperature, top_k, top_p, rep, device, seed):
        calls.append((len(starts), length, seed))
        rows = []
        for s in starts:
            assert s["targets"].tolist() == [1] and "rating" not in s
            tag = 3 + int(s["topic_ids"][0]) % 2  # 'a' for even items, 'b' for odd ones
            rows.append([1, tag, tag, 2, 1, tag, 2, 5, 0])
        return rows

    data = [{"topic_ids": np.array([i]), "rating": np.array(3)} for i in range(5)]
    path = tmp_path / "samples.txt"
    out = generate_samples(Model(), data, Tok(), n_samples=3, rows_per_call=4, save_path=str(path),
                           _sampler=fake_sampler)
    assert len(out) == 15 and [c[0] for c in calls] == [4, 4, 4, 3] and all(c[1] == 7 for c in calls)
    assert out[:3] == ["aa，a"] * 3 and out[3:6] == ["bb，b"] * 3
    assert path.read_text(encoding="utf-8").splitlines() == out


def test_ddp_graph_segments_cover_every_stage_once():
    from mmtg_b200.graph import segment_bounds
    for nstage in (15, 5, 3):
        for group in (1, 2, 4, 7, 100):
            b = segment_bounds(nstage, group)
            flat = [s for s0, s1 in b for s in range(s0, s1)]
            assert flat == list(range(nstage)), (nstage, group, b)
            assert b[-1] == (nstage - 1, nstage)  # the encoder-side stage is its own segment
            assert all(s1 - s0 <= max(1, group) for s0, s1 in b)
    assert segment_bounds(15, 4) == [(0, 4), (4, 8), (8, 12), (12, 14), (14, 15)]


def test_package_exports_the_drop_in_surface():
    """`from mmtg_b200 import MMTG, MyLoss, ...` is the 3-import switch INTEGRATION.md shows."""
    import mmtg_b200
    for name in ("MMTG", "MyLoss", "sample_sequence", "top_k_top_p_filtering", "model_cfgs", "data_config",
                 "FusedAdamW", "generate_samples"):
        assert hasattr(mmtg_b200, name), name
