"""Data-parallel gradient exchange (replaces torch.nn.DataParallel, src/train.py:112-114).

One process per GPU, identical replicas, batch sharded by rank. The backward engine runs in
stages (lm_head, block 11 .. block 0, embeddings/encoder); as soon as a stage has produced the
gradients of one contiguous bucket of the flat gradient buffer (a GPT-2 block = 7.1 M params =
28 MB fp32) that bucket is all-reduced (NCCL over NVLink/NVSwitch through torch.distributed) on a
side stream while the next stage computes. There is no parameter broadcast and no logits gather.
Equal per-rank batch sizes -> average; see DESIGN.md for the ragged (curriculum-filtered) case.
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
        nl = nstage - 2
        if 1 <= stage <= nl:
            lo, hi = model.layer_bucket(nl - stage)
            self._reduce(G[lo:hi])
        elif stage == nstage - 1:
            lo, hi = model.tail_bucket()
            self._reduce(G[lo:hi])

    def finish(self, model):
        if self.world > 1 and self._cuda and model._flat[2].is_cuda:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
