"""Greedy B=64 x 220 generation, CUDA-graph replay: tokens/s and us per position (best of 3)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mmtg_b200 import synth
from mmtg_b200.configs import data_config, model_cfgs
from mmtg_b200.generate import sample_sequence_batch
from mmtg_b200.model import MMTG
B, LENGTH = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 220
model = MMTG(model_cfgs, data_config(), 13317, train_flag=False, token_table=synth.make_token_table())
model.load_state_dict(synth.make_state_dict(0)); model.to("cuda:0")
batch = synth.make_batch(B, seed=1234)
starts = {k: v for k, v in batch.items() if k != "rating"}
starts["targets"] = np.ones((B, 1), np.int64)
ts = []
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rows = sample_sequence_batch(model, starts, LENGTH, device="cuda", use_cuda_graph=True, temperature=1.0, top_k=1, top_p=0.0, repitition_penalty=1.0)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
t = min(ts)
print(f"tokens/s {B*LENGTH/t:.0f}  us/position {t/LENGTH*1e6:.1f}  ids_sum {sum(sum(r) for r in rows)}")
