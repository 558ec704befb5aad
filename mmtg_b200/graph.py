"""CUDA-graphed training step for fixed batch shapes.

The train step issues ~380 kernel launches; on a busy host the Python/driver launch path, not the
GPU, sets the step time. `GraphedTrainStep` captures forward + MyLoss + alpha*KL + backward
(+ gradient all-reduce) + clip + AdamW once (through `MMTG.fused_train_step`, the autograd-free
driver of the same engine entry points) and replays it; inputs are copied into static device
buffers before each replay, learning-rate / step-count state lives on the device
(mmtg_b200.optim.FusedAdamW). Shapes must not change between calls (curriculum-filtered batches
of another size need their own instance).
"""
from __future__ import annotations

import torch


class GraphedTrainStep:
    def __init__(self, model, criterion, optimizer, example_batch, alpha=0.2, stage=3, warmup=3):
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.alpha, self.stage = alpha, stage
        self.static = {k: v.detach().clone() for k, v in example_batch.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.total = self._eager()

    def _eager(self):
        # engine-driven step (no torch.autograd inside the capture): same kernels, same C-ABI
        total, _loss, _kl = self.model.fused_train_step(self.static, self.stage, self.alpha)
        self.optimizer.step()
        self.optimizer.zero_grad()
        return total

    def __call__(self, batch):
        """batch: dict of tensors (device, or pinned host — copied asynchronously). Returns the
        total loss as a device scalar (valid until the next call)."""
        for k, dst in self.static.items():
            src = batch[k]
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.total
