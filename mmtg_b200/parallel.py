"""Data-parallel gradient exchange (replaces torch.nn.DataParallel, src/train.py:112-114).

One process per GPU, identical replicas, batch sharded by rank. The backward engine runs in
stages (lm_head, block 11 .. block 0, embeddings + projector, encoder side); as soon as a stage has produced the
gradients of one contiguous bucket of the flat gradient buffer (a GPT-2 block = 7.1 M params =
28 MB fp32) that bucket is all-reduced (NCCL over NVLink/NVSwitch through torch.distributed) on a
side stream while the next stage computes. There is no parameter broadcast and no logits gather.
Equal per-rank batch sizes -> average (default). Ragged per-rank batches (curriculum stages 1-2
filter rows by rating, src/train.py:178-183): build `GradSync(average=False)` and scale each
rank's loss gradient by `ragged_batch_scale(B_local)` = B_local / B_global
(`MMTG.fused_train_step(..., grad_scale=...)`), so the SUM all-reduce reproduces the
single-process gradient of the concatenated batch.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, process_group=None, average=True):
        self.group = process_group
        self.average = average
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._cuda = torch.cuda.is_available()
        self.comm_stream = torch.cuda.Stream() if self._cuda else None
        self.bytes_reduced = 0

    def _reduce(self, t):
        if self.world == 1 or t.numel() == 0:
            return
        self.bytes_reduced += t.numel() * t.element_size()
        if t.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                if self.average:
                    dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
                else:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        else:  # gloo (CPU tests of the bucketing logic): no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                t.div_(self.world)

    def after_stage(self, model, stage, nstage):
        G = model._flat[2]
        nl = nstage - 3
        if 1 <= stage <= nl:
            lo, hi = model.layer_bucket(nl - stage)
            self._reduce(G[lo:hi])
        elif stage == nl + 1:  # projector, wpe, ln_f, tied wte: reduced while the encoder side runs
            lo, hi = model.tail_buckets()[0]
            self._reduce(G[lo:hi])
        elif stage == nl + 2:  # encoder + multi-modal attention
            lo, hi = model.tail_buckets()[1]
            self._reduce(G[lo:hi])

    def finish(self, model):
        if self.world > 1 and self._cuda and model._flat[2].is_cuda:
            torch.cuda.current_stream().wait_stream(self.comm_stream)


def ragged_batch_scale(b_local: int, device=None, group=None) -> float:
    """B_local / B_global for this step (one tiny integer all-reduce)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 1.0
    t = torch.tensor([float(b_local)], device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(b_local) / float(t.item())
