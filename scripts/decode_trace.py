"""Per-phase timeline of the decode megakernel (CTA 0 globaltimer stamps) at the last position."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mmtg_b200 import _lib, synth
from mmtg_b200.configs import data_config, model_cfgs
from mmtg_b200.generate import sample_sequence_batch
from mmtg_b200.model import MMTG

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
model = MMTG(model_cfgs, data_config(), 13317, train_flag=False, token_table=synth.make_token_table())
model.load_state_dict(synth.make_state_dict(0))
model.to(dev)
batch = synth.make_batch(B, seed=1234)
starts = {k: v for k, v in batch.items() if k != "rating"}
starts["targets"] = np.ones((B, 1), np.int64)
buf = torch.zeros(512, dtype=torch.int64, device=dev)
lib = _lib.lib()
lib.mmtg_decode_set_trace.argtypes = [C.c_void_p]
lib.mmtg_decode_set_trace(C.c_void_p(buf.data_ptr()))
sample_sequence_batch(model, starts, 220, device="cuda", use_cuda_graph=False, temperature=1.0, top_k=1, top_p=0.0,
                      repitition_penalty=1.0)
torch.cuda.synchronize()
t = buf.cpu().numpy()
# stamps: kernel start, after the prologue barrier, after each of the 5 phase barriers of the 12 blocks, kernel end
n = 2 + 5 * 12 + 1
d = np.diff(t[:n]) / 1e3
print("total us %.1f  prologue %.2f" % (d.sum(), d[0]))
for ph in "ABCDE":
    v = [d[1 + 5 * l + "ABCDE".index(ph)] for l in range(12)]
    print(ph, "mean %.2f us  min %.2f max %.2f" % (np.mean(v), np.min(v), np.max(v)))
print("F %.2f" % d[61])
names = ["start", "acts landed", "weights landed", "mma done", "cluster sync 1", "finalize done", "cluster sync 2", "grid arrive", "window done", "grid wait done"]
for label, o in (("A", 80), ("C", 90), ("D", 100), ("E", 110)):
    v = t[o:o + 10].astype(np.float64)
    base = v[0]
    print(label, "(last block, CTA 0) us from phase start:", {names[i]: round((v[i] - base) / 1e3, 2) for i in range(10) if v[i] > 0})
v = t[120:123].astype(np.float64)
print("B (last block, CTA 0 thread 0): attention work %.2f us, arrive %.2f us" % ((v[1] - v[0]) / 1e3, (v[2] - v[1]) / 1e3))
