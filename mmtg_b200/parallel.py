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

Wire format. fp32 by default (N-rank gradients equal the single-rank ones to 4e-7,
tests/ddp_worker.py). `GradSync(grad_dtype=torch.bfloat16)` / MMTG_DDP_GRAD_DTYPE=bf16 send each
bucket as bf16 (cast on the communication stream, all-reduce of half the bytes, cast back into the
fp32 gradient buffer; 2^-9 relative rounding per element and rank, measured 2.7e-3 worst tensor).
MEASURED (round 2, B200, 32 samples per GPU): 8 GPUs 9.10 ms (bf16) vs 9.20 ms (fp32) per step,
2 GPUs 8.89 ms vs 8.76 ms - the two extra cast passes over the gradient buffer (1.3 GB of HBM
traffic per step) cost about what the shorter collective saves, so it stays opt-in. Also
measured at 8 GPUs: capping NCCL at 16 / 8 channels lengthens the step to 9.80 / 11.64 ms (the
collective's duration, not the SMs it occupies, is what is exposed), and dynamic GEMM tile
scheduling (MMTG_GEMM_CLC=1) does not shorten it (9.32 vs 9.25 ms).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, process_group=None, average=True, grad_dtype=None):
        self.group = process_group
        self.average = average
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._cuda = torch.cuda.is_available()
        self.comm_stream = torch.cuda.Stream() if self._cuda else None
        self.bytes_reduced = 0
        if grad_dtype is None:
            grad_dtype = torch.float32 if os.environ.get("MMTG_DDP_GRAD_DTYPE", "fp32").lower() in ("fp32", "float32") \
                else torch.bfloat16
        self.grad_dtype = grad_dtype
        self._wire = None  # bf16 staging buffer, same offsets as the flat gradient buffer
        self._wte_saved, self._wte_evt = None, None  # early all-reduce of the tied wte gradient (after_stage)

    def _reduce(self, t):
        if self.world == 1 or t.numel() == 0:
            return
        self.bytes_reduced += t.numel() * t.element_size()
        if t.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            op = dist.ReduceOp.AVG if self.average else dist.ReduceOp.SUM
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                if self.grad_dtype == torch.float32 or t.dtype != torch.float32:
                    dist.all_reduce(t, op=op, group=self.group)
                else:
                    w = self._wire_view(t)
                    self.bytes_reduced -= t.numel() * (t.element_size() - w.element_size())
                    w.copy_(t)  # fp32 -> bf16 on the communication stream
                    dist.all_reduce(w, op=op, group=self.group)
                    t.copy_(w)
        else:  # gloo (CPU tests of the bucketing logic): no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                t.div_(self.world)

    def _on_comm(self, t):
        """Context for small follow-up ops that must run after the collective just issued for `t`."""
        import contextlib
        return torch.cuda.stream(self.comm_stream) if t.is_cuda else contextlib.nullcontext()

    def _wire_view(self, t):
        """bf16 staging slice for the gradient slice `t` (one buffer, allocated once: the caching
        allocator must not hand out memory that a CUDA-graph pool owns)."""
        n = t.untyped_storage().nbytes() // t.element_size()
        if self._wire is None or self._wire.numel() != n or self._wire.device != t.device:
            self._wire = torch.empty(n, dtype=self.grad_dtype, device=t.device)
        o = t.storage_offset()
        return self._wire[o:o + t.numel()]

    # Rows of the tied wte that the embedding backward (token-TYPE embeddings, stage NL+1) can touch.
    # The gradient of wte is the lm_head weight gradient (all 13,317 rows, final after stage 0) plus
    # the type-embedding rows; type ids are sentence indices (src/MyDataset.py: <= max_seq_length /
    # sent_len + 1), far below this bound.
    WTE_TYPE_ROWS = 64

    def _type_rows_ok(self, model):
        """The early wte all-reduce is only valid when every token-type id the dataset can produce lies in the
        first WTE_TYPE_ROWS rows (src/MyDataset.py: type ids are sentence indices 1 .. max_seq_length // sent_len
        + 1); otherwise the tied gradient goes with the tail bucket as before."""
        try:
            dc = model.data_config
            return dc["max_seq_length"] // (dc["max_sent_length"] + 2) + 2 <= self.WTE_TYPE_ROWS
        except Exception:
            return False

    def before_stages(self, model, s0, s1, nstage):
        """Called before backward stages [s0, s1) are issued on the current stream."""
        nl = nstage - 3
        if self._wte_saved is not None and s0 <= nl + 1 < s1 and self._wte_evt is not None:
            # the embedding stage adds into wte rows the early all-reduce has already saved and cleared
            torch.cuda.current_stream().wait_event(self._wte_evt)

    def after_stage(self, model, stage, nstage):
        G = model._flat[2]
        nl = nstage - 3
        early_wte = self.world > 1 and hasattr(model, "wte_range") and self._type_rows_ok(model)
        if stage == 0 and early_wte:
            # The tied wte / lm_head gradient (41 MB, the largest single tensor) is all-reduced NOW, under the
            # whole decoder backward, instead of in the exposed tail: by linearity the type-embedding rows that
            # stage NL+1 still adds are reduced separately - the reduced head part of those rows is set aside and
            # the rows are cleared (so the tail bucket carries only the embedding contribution), and added back
            # after the tail all-reduce.
            lo, hi, E = model.wte_range()
            rows = min(self.WTE_TYPE_ROWS, (hi - lo) // E)
            self._reduce(G[lo:hi])
            with self._on_comm(G):
                head = G[lo:lo + rows * E]
                if self._wte_saved is None or self._wte_saved.numel() != head.numel() or self._wte_saved.device != head.device:
                    self._wte_saved = torch.empty_like(head)
                self._wte_saved.copy_(head)
                head.zero_()
                if G.is_cuda:
                    self._wte_evt = torch.cuda.Event()
                    self._wte_evt.record()
        elif 1 <= stage <= nl:
            lo, hi = model.layer_bucket(nl - stage)
            self._reduce(G[lo:hi])
        elif stage == nl + 1:  # projector, wpe, ln_f (+ the type rows of wte): reduced while the encoder side runs
            lo, hi = model.tail_buckets()[0]
            if early_wte and self._wte_saved is not None:
                wlo, _whi, E = model.wte_range()
                hi = wlo + self._wte_saved.numel()
                self._reduce(G[lo:hi])
                with self._on_comm(G):
                    G[wlo:hi].add_(self._wte_saved)
            else:
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
