"""KV-cached generation throughput (BASELINE.json configs[3]: batch 64, 220 positions, 1 GPU).
tokens/s = B * 220 / wall; HBM roofline: 193.2 MB of bf16 weights per step + KV."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mmtg_b200 import _lib, synth
from mmtg_b200.configs import data_config, model_cfgs
from mmtg_b200.generate import sample_sequence_batch
from mmtg_b200.model import MMTG

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
LENGTH = 220
dev = torch.device("cuda:0")
table = synth.make_token_table()
model = MMTG(model_cfgs, data_config(), 13317, train_flag=False, token_table=table)
model.load_state_dict(synth.make_state_dict(0))
model.to(dev)
batch = synth.make_batch(B, seed=1234)
starts = {k: v for k, v in batch.items() if k != "rating"}
starts["targets"] = np.ones((B, 1), np.int64)
out = {}
for name, kw in (("greedy", dict(temperature=1.0, top_k=1, top_p=0.0, repitition_penalty=1.0)),
                 ("topk10_p0.7", dict(temperature=1.1, top_k=10, top_p=0.7, repitition_penalty=1.5))):
    for graph in (True, False):
        ts = []
        for rep in range(3):
            if rep == 2 and name == "greedy" and not graph:
                torch.cuda.profiler.start()  # ncu --profile-from-start off: one eager generation
            torch.cuda.synchronize()
            l0 = _lib.launch_count()
            t0 = time.perf_counter()
            rows = sample_sequence_batch(model, starts, LENGTH, device="cuda", use_cuda_graph=graph, **kw)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
            if rep == 2 and name == "greedy" and not graph:
                torch.cuda.profiler.stop()
        t = min(ts)
        weights_mb = 193.2
        kv_gb = sum(36864 * (15 + j) for j in range(LENGTH)) * B / 1e9
        hbm_floor = (LENGTH * weights_mb / 1e3 + kv_gb) / 6548.2
        out[f"{name}_{'graph' if graph else 'eager'}"] = {
            "tokens_per_s": B * LENGTH / t, "wall_s": t, "ms_per_step": t / LENGTH * 1e3,
            "hbm_floor_s": hbm_floor, "x_of_roofline": t / hbm_floor, "launches": _lib.launch_count() - l0,
            "len_returned": len(rows[0])}
print(json.dumps({"metric": "decode tokens/s", "batch": B, "length": LENGTH, "results": out}))
