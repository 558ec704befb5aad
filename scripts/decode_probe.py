"""Short eager decode run for ncu launch lists (B=64, 40 positions)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mmtg_b200 import synth
from mmtg_b200.configs import data_config, model_cfgs
from mmtg_b200.generate import sample_sequence_batch
from mmtg_b200.model import MMTG
B = 64
dev = torch.device("cuda:0")
model = MMTG(model_cfgs, data_config(), 13317, train_flag=False, token_table=synth.make_token_table())
model.load_state_dict(synth.make_state_dict(0)); model.to(dev)
batch = synth.make_batch(B, seed=1234)
starts = {k: v for k, v in batch.items() if k != "rating"}
starts["targets"] = np.ones((B, 1), np.int64)
rows = sample_sequence_batch(model, starts, 40, device="cuda", use_cuda_graph=False, temperature=1.1, top_k=10, top_p=0.7, repitition_penalty=1.5)
torch.cuda.synchronize(); print(len(rows), len(rows[0]))
