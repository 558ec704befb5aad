"""CUDA-graphed training step for fixed batch shapes.

The train step issues ~380 kernel launches; on a busy host the Python/driver launch path, not the
GPU, sets the step time. `GraphedTrainStep` captures the step once and replays it. It drives the
engine through `MMTG.fused_forward_loss` / `backward_stages` (the autograd-free drivers of the
same C-ABI entry points); inputs are copied into static device buffers before each replay and
learning-rate / step-count state lives on the device (mmtg_b200.optim.FusedAdamW).

Single GPU: one graph (forward + MyLoss + alpha*KL + backward + clip + AdamW).
Data parallel (`model.grad_sync` set): one graph for forward+loss, one per backward stage and one
for the optimizer, sharing a memory pool; the bucketed NCCL all-reduces stay EAGER and are issued
between the stage replays on GradSync's side stream (no collectives inside a capture).
Shapes must not change between calls (a curriculum-filtered batch of another size needs its own
instance).
"""
from __future__ import annotations

import torch


def segment_bounds(nstage: int, group: int):
    """[s0, s1) backward-stage ranges of the captured graph segments of a data-parallel step:
    `group` stages per segment, except that the last stage (the encoder side) is always a segment
    of its own — the projector / tied-wte gradient bucket that is final before it is all-reduced
    while it runs."""
    group = max(1, int(group))
    bounds = [(s0, min(nstage - 1, s0 + group)) for s0 in range(0, nstage - 1, group)]
    bounds.append((nstage - 1, nstage))
    return bounds


class GraphedTrainStep:
    def __init__(self, model, criterion, optimizer, example_batch, alpha=0.2, stage=3, warmup=3, grad_scale=1.0):
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.alpha, self.stage = alpha, stage
        self.grad_scale = grad_scale  # B_local / B_global of a ragged data-parallel step (parallel.py)
        self.static = {k: v.detach().clone() for k, v in example_batch.items()}
        self.sync = model.grad_sync
        # capture on a HIGH-priority stream: the engine's weight-gradient side stream has the lowest
        # priority, so the backward chain's kernels win SMs as they free up (priorities are kept
        # in the captured kernel nodes)
        side = torch.cuda.Stream(priority=-1)
        self.cap_stream = side
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if self.sync is None:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.cap_stream):
                self.total = self._eager()
            return
        # segmented capture: collectives are issued eagerly between the segments
        self.g_fwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fwd, stream=self.cap_stream):
            self.step, self.total, _, _ = model.fused_forward_loss(self.static, stage, alpha, grad_scale)
        pool = self.g_fwd.pool()
        self.nstage = self.step.dims.NL + 3
        # backward stages per captured segment: 1 = an all-reduce can start after every block;
        # larger groups trade overlap granularity for fewer graph launches / stream hand-offs
        # (measured ms/step at 2 GPUs: 9.43 / 9.30 / 9.20 for 1 / 2 / 4; at 8 GPUs 9.72 / 9.63 for 2 / 4)
        import os
        group = max(1, int(os.environ.get("MMTG_DDP_STAGE_GROUP", "4")))
        self.g_stage = []  # (graph, first stage, one-past-last stage)
        for s0, s1 in segment_bounds(self.nstage, group):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, stream=self.cap_stream):
                model.backward_stages(self.step, s0, s1)
            self.g_stage.append((g, s0, s1))
        self.g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_opt, pool=pool, stream=self.cap_stream):
            optimizer.step()
            optimizer.zero_grad()

    def _eager(self):
        total, _loss, _kl = self.model.fused_train_step(self.static, self.stage, self.alpha, self.grad_scale)
        self.optimizer.step()
        self.optimizer.zero_grad()
        return total

    def prefetch(self, batch):
        """Double-buffered input staging: start copying the NEXT step's batch
        (pinned host tensors) into device staging buffers on a copy stream while the current
        step runs; the following `__call__(batch)` with the same dict only pays a device-to-device
        copy into the graph's static inputs."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
            self._staging = {k: torch.empty_like(v) for k, v in self.static.items()}
            self._staged_evt, self._consumed_evt = torch.cuda.Event(), None
        with torch.cuda.stream(self._copy_stream):
            if self._consumed_evt is not None:  # the previous staging contents have been consumed
                self._copy_stream.wait_event(self._consumed_evt)
            for k, dst in self._staging.items():
                dst.copy_(batch[k], non_blocking=True)
            self._staged_evt.record()
        self._staged = batch

    def __call__(self, batch):
        """batch: dict of tensors (device, or pinned host — copied asynchronously; or the dict
        last given to `prefetch`). Returns the total loss as a device scalar (valid until the
        next call)."""
        if getattr(self, "_staged", None) is batch:
            cur = torch.cuda.current_stream()
            cur.wait_event(self._staged_evt)
            for k, dst in self.static.items():
                dst.copy_(self._staging[k], non_blocking=True)
            self._consumed_evt = torch.cuda.Event()
            self._consumed_evt.record()
            self._staged = None
        else:
            for k, dst in self.static.items():
                src = batch[k]
                if src is not dst:
                    dst.copy_(src, non_blocking=True)
        if hasattr(self.optimizer, "sync_lr"):
            self.optimizer.sync_lr()  # an LR scheduler's param_groups['lr'] -> the device copy the graph reads
        if self.sync is None:
            self.graph.replay()
            return self.total
        self.g_fwd.replay()
        for g, s0, s1 in self.g_stage:
            self.sync.before_stages(self.model, s0, s1, self.nstage)
            g.replay()
            for s in range(s0, s1):
                self.sync.after_stage(self.model, s, self.nstage)
        self.sync.finish(self.model)
        self.g_opt.replay()
        return self.total
