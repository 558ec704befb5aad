"""Writes tests/golden/bench_b32_step.json: the CPU oracle's losses for bench.py's exact rank-0
inputs (batch 32, seed 1234, state-dict seed 0, stage 3, alpha 0.2, no dropout). bench.py checks
its first step against these numbers without importing oracle/ (the oracle is test infrastructure).
Also records the same quantities for the extended lengths bench.py --max-sent-length accepts.
Run here (CPU): python scripts/make_bench_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from mmtg_b200 import synth
from mmtg_b200.configs import data_config
from oracle import mmtg_oracle as O

torch.set_num_threads(os.cpu_count() or 1)
table = torch.from_numpy(synth.make_token_table())
sd = synth.make_state_dict(0)
host = synth.batch_to_torch(synth.make_batch(32, seed=1234))
with torch.no_grad():
    hf, kl, logits = O.mmtg_forward(sd, table, host, data_config(), True)
    out = {"batch": 32, "seed": 1234, "state_dict_seed": 0, "alpha": 0.2, "hf_loss": float(hf), "kl": float(kl)}
    for s in (1, 2, 3):
        out[f"myloss_stage{s}"] = float(O.my_loss(logits, host["targets"], host["rating"], s))
    out["total"] = out["myloss_stage3"] + 0.2 * out["kl"]
    out["source"] = "oracle/mmtg_oracle.py (fp32, CPU), pinned to the unmodified reference by tests/test_oracle_vs_reference.py"
path = os.path.join(ROOT, "tests", "golden", "bench_b32_step.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print(path, out)
