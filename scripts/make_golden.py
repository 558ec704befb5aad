"""Generates tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference/src) in
the build container (see oracle/ref_import.py). Inputs/weights are regenerated from seeds by
mmtg_b200.synth, so only outputs are stored (sub-sampled: full logits are 25 MB).

    python scripts/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmtg_b200 import synth  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ALPHA = 0.2  # src/train.sh:15


class _Tok:  # the only tokenizer call on the path: src/generate.py:133-136
    _m = {"[#START#]": 1, "[#EOS#]": 2, "[UNK]": 100, "[SEP]": 102}

    def convert_tokens_to_ids(self, t):
        return self._m[t]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    table = synth.make_token_table()
    sd = synth.make_state_dict(0)
    model, crit, gen, dc, cfgs = ref_import.load_reference(table, sd)

    # ---- config 1: forward + loss, batch 2 (BASELINE.json configs[0]) ----
    B = 2
    ratings = np.array([5, 2])
    batch = synth.batch_to_torch(synth.make_batch(B, seed=1234, ratings=ratings))
    res = {}
    for p in model.parameters():
        p.grad = None
    hf_loss, kl, logits = model(batch)
    res["hf_loss"] = hf_loss.item()
    res["kl"] = kl.item()
    lg = logits.detach()
    res["logits_sub"] = lg[:, ::5, ::97].numpy()
    res["logits_rows"] = lg[:, [0, 14, 15, 100, 235], :].numpy()
    res["logits_rowsum"] = lg.sum(-1).numpy()
    res["logits_argmax"] = lg.argmax(-1).numpy()
    for stage in (1, 2, 3):
        res[f"myloss_stage{stage}"] = crit(lg, batch["targets"], batch["rating"], stage).item()
    # gradient oracle: the restated train step of src/train.py:188-193, stage 3
    loss = crit(logits, batch["targets"], batch["rating"], 3)
    total = loss.mean() + ALPHA * kl.mean()
    total.backward()
    res["total_loss"] = total.item()
    names, norms, heads = [], [], []
    for n, p in model.named_parameters():
        g = p.grad.detach().flatten()
        names.append(n)
        norms.append(g.norm().item())
        idx = torch.linspace(0, g.numel() - 1, 32).long()
        heads.append(g[idx].numpy())
    res["grad_names"] = np.array(names)
    res["grad_norms"] = np.array(norms)
    res["grad_samples"] = np.stack(heads)
    np.savez_compressed(os.path.join(OUT, "c1_forward_loss_b2.npz"), **res)
    print("c1:", res["hf_loss"], res["kl"], res["myloss_stage3"], res["total_loss"])

    # ---- negative-sample ratio sweep on B=4 (MyLoss only needs logits/targets/ratings) ----
    b4 = synth.batch_to_torch(synth.make_batch(4, seed=77))
    with torch.no_grad():
        _, kl4, lg4 = model(b4)
        sweep = {"kl": kl4.item(), "logits_rowsum": lg4.sum(-1).numpy()}
        for r in ([5, 5, 5, 5], [5, 4, 3, 1], [1, 2, 1, 3], [4, 4, 2, 5]):
            for stage in (1, 2, 3):
                key = "r" + "".join(map(str, r)) + f"_s{stage}"
                sweep[key] = crit(lg4, b4["targets"], torch.tensor(r), stage).item()
    np.savez_compressed(os.path.join(OUT, "loss_sweep_b4.npz"), **sweep)

    # ---- generation: greedy + top-k/top-p filtering goldens ----
    model.train_flag = False
    one = synth.make_batch(1, seed=99)
    start = {k: v[0] for k, v in one.items() if k != "rating"}
    start["targets"] = np.asarray([1])
    rows = []
    orig_forward = model.forward

    def rec_forward(inputs):
        out = orig_forward(inputs)
        rows.append(out[2][0, -1, :].detach().clone().numpy())
        return out

    model.forward = rec_forward
    LEN = 48
    ids = gen.sample_sequence(model, dict(start), LEN, _Tok(), temperature=1.0, top_k=1, top_p=0.0,
                              repitition_penalty=1.0, device="cpu")
    step_logits = np.stack(rows)
    g = {"greedy_ids": np.array(ids), "greedy_step_logits_sub": step_logits[:, ::13],
         "greedy_step_top2": np.sort(step_logits, -1)[:, -2:], "length": LEN}
    # with the CLI preset (src/generate.sh:9-12) the draw is random; pin the filtered distribution
    rows.clear()
    torch.manual_seed(5)
    ids2 = gen.sample_sequence(model, dict(start), 24, _Tok(), temperature=1.1, top_k=10, top_p=0.7,
                               repitition_penalty=1.5, device="cpu")
    g["sampled_ids_seed5"] = np.array(ids2)
    model.forward = orig_forward
    # top_k_top_p_filtering known-answer vectors
    rng = np.random.default_rng(3)
    for i, (k, p) in enumerate([(10, 0.7), (0, 0.9), (5, 0.0), (1, 0.0), (30, 0.3)]):
        x = torch.from_numpy(rng.standard_normal(13317).astype(np.float32) * 2)
        g[f"filt_in_{i}"] = x.numpy().copy()
        y = gen.top_k_top_p_filtering(x.clone(), top_k=k, top_p=p)
        g[f"filt_keep_{i}"] = torch.isfinite(y).nonzero().flatten().numpy()
        g[f"filt_kp_{i}"] = np.array([k, p])
    np.savez_compressed(os.path.join(OUT, "generate_b1.npz"), **g)
    print("greedy ids:", ids[:30])


if __name__ == "__main__":
    main()
