"""Top-2 logit margins of the CPU oracle along the greedy sequences of the generation test rows:
shows where run-to-run differences of the fused decode step (atomic summation order) can flip a token."""
import sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmtg_b200 import synth
from mmtg_b200.configs import data_config
from oracle import mmtg_oracle as O
torch.set_num_threads(8)
table = torch.from_numpy(synth.make_token_table())
sd = synth.make_state_dict(0)
for seed in (99, 7, 21):
    one = synth.make_batch(1, seed=seed)
    start = {k: v[0] for k, v in one.items() if k != "rating"}
    start["targets"] = np.asarray([1])
    ids, logits = O.sample_sequence(sd, table, start, 60, data_config(), temperature=1.0, top_k=1, top_p=0.0,
                                    repitition_penalty=1.0, return_logits=True)
    gaps = []
    for lg in logits:
        t = torch.topk(lg, 2).values
        gaps.append((t[0] - t[1]).item())
    gaps = np.array(gaps)
    small = [(i, round(g, 4)) for i, g in enumerate(gaps) if g < 0.03]
    print("seed", seed, "n_model_steps", len(gaps), "min gap %.4f" % gaps.min(), "steps with gap<0.03:", small[:12])
    print("   ids[:20]", ids[:20])
