"""Run-to-run reproducibility of the decode step: generates the same rows N times with the fused
persistent-kernel step (atomic split-K sums) and with the per-op step, and reports the largest
difference of the step logits between runs while the generated histories still agree, plus the
first position where the greedy ids of two runs part (and the top-2 margin there)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mmtg_b200 import synth
from mmtg_b200.configs import data_config, model_cfgs
from mmtg_b200.generate import sample_sequence_batch
from mmtg_b200.model import MMTG

N = int(sys.argv[1]) if len(sys.argv) > 1 else 6
LENGTH = 60
model = MMTG(model_cfgs, data_config(), 13317, train_flag=False, token_table=synth.make_token_table())
model.load_state_dict(synth.make_state_dict(0))
model.to("cuda")
starts = []
for seed in (99, 7, 21):
    one = synth.make_batch(1, seed=seed)
    s = {k: v[0] for k, v in one.items() if k != "rating"}
    s["targets"] = np.asarray([1])
    starts.append(s)
kw = dict(temperature=1.0, top_k=1, top_p=0.0, repitition_penalty=1.0, device="cuda", return_step_logits=True)
for mode in ("1", "0"):
    os.environ["MMTG_DECODE_MEGA"] = mode
    runs = [sample_sequence_batch(model, starts, LENGTH, **kw) for _ in range(N)]
    ids0, lg0 = runs[0]
    worst, parts = 0.0, []
    for ids, lg in runs[1:]:
        for r in range(len(starts)):
            same = 0
            while same < len(ids[r]) and ids[r][same] == ids0[r][same]:
                same += 1
            for k in range(min(same, len(lg), len(lg0))):
                worst = max(worst, (lg[k][r] - lg0[k][r]).abs().max().item())
            if same < len(ids[r]):
                top = torch.topk(lg0[same - 1][r].float(), 2).values
                parts.append((r, same, round((top[0] - top[1]).item(), 4)))
    print(f"{'fused' if mode == '1' else 'per-op'} step: {N} runs, max |dlogit| between runs on equal histories "
          f"{worst:.2e}; (row, first differing position, top-2 margin there): {parts}")
