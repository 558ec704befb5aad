import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from mmtg_b200 import synth
from mmtg_b200.configs import data_config, model_cfgs
from mmtg_b200.generate import sample_sequence_batch
from mmtg_b200.model import MMTG
B = int(sys.argv[1])
model = MMTG(model_cfgs, data_config(), 13317, train_flag=False, token_table=synth.make_token_table())
model.load_state_dict(synth.make_state_dict(0)); model.to("cuda")
batch = synth.make_batch(B, seed=1234)
starts = {k: v for k, v in batch.items() if k != "rating"}
starts["targets"] = np.ones((B, 1), np.int64)
kw = dict(temperature=1.0, top_k=1, top_p=0.0, repitition_penalty=1.0, device="cuda")
for L in (220, 20):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sample_sequence_batch(model, starts, L, **kw)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
    print("B", B, "len", L, "wall ms %.2f" % (t * 1e3))
