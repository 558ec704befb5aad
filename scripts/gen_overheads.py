"""Time the per-call setup pieces of generation (prefix load, weight folding) with CUDA events."""
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

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model = MMTG(model_cfgs, data_config(), 13317, train_flag=False, token_table=synth.make_token_table())
model.load_state_dict(synth.make_state_dict(0))
model.to("cuda")
batch = synth.make_batch(B, seed=1234)
starts = {k: v for k, v in batch.items() if k != "rating"}
starts["targets"] = np.ones((B, 1), np.int64)
sample_sequence_batch(model, starts, 220, device="cuda", top_k=1)
ses = next(iter(model._decode_sessions.values()))
lib = _lib.lib()
st = C.c_void_p(_lib.stream_ptr())
cm = ses.cm


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("fold ms", timed(lambda: lib.mmtg_decode_fold_weights(C.byref(cm), ses.Lmax, C.c_void_p(ses.dws.data_ptr()), st)))
import time
t0 = time.perf_counter()
for _ in range(5):
    lib.mmtg_decode_fold_weights(C.byref(cm), ses.Lmax, C.c_void_p(ses.dws.data_ptr()), st)
print("fold host-side issue ms", (time.perf_counter() - t0) / 5 * 1e3)
torch.cuda.synchronize()
