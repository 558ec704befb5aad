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
n = 62
d = np.diff(t[:n]) / 1e3
print("total us", d.sum())
for ph in "ABCDE":
    v = [d[5 * l + "ABCDE".index(ph)] for l in range(12)]
    print(ph, "mean %.2f us  min %.2f max %.2f" % (np.mean(v), np.min(v), np.max(v)))
print("F %.2f" % d[60])
a0 = t[5 * 11]  # stamp at the start of the last block's phase A
sub = t[64:70]
print("phase A (last block) sub-stamps us from phase start [stats, side+sync, stage, wait+sync, mma+red, prefetch]:",
      [round(float(x - a0) / 1e3, 2) for x in sub], "end", round(float(t[5 * 11 + 1] - a0) / 1e3, 2))
arrA = (t[80:80 + 148] - a0) / 1e3
arrB = (t[240:240 + 148] - t[5 * 11 + 1]) / 1e3
print("arrival at A->B barrier (us after CTA0 phase start): min %.2f med %.2f max %.2f argmax %d" % (arrA.min(), np.median(arrA), arrA.max(), arrA.argmax()))
print("  sorted tail:", np.sort(arrA)[-8:].round(2), "ctas", np.argsort(arrA)[-8:])
print("arrival at B->C barrier (us after CTA0 B start): min %.2f med %.2f max %.2f argmax %d" % (arrB.min(), np.median(arrB), arrB.max(), arrB.argmax()))
print("  sorted tail:", np.sort(arrB)[-8:].round(2), "ctas", np.argsort(arrB)[-8:])
f0 = t[60]
print("phase F sub-stamps us [staged, unit0, unit1, unit2]:", [round(float(x - f0) / 1e3, 2) for x in t[72:76]])
arr0 = (t[400:400 + 148] - a0) / 1e3
print("reached A->B arrive: min %.2f med %.2f max %.2f" % (arr0.min(), np.median(arr0), arr0.max()), " tail", np.sort(arr0)[-6:].round(2))
