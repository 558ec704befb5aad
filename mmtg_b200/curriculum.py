"""Curriculum batch assembly (restates src/train.py:154-186).

Stage schedule: epoch < curriculums[0] -> stage 1 (only very negative / very positive samples,
rating < 2 or > 4); epoch < curriculums[1] -> stage 2 (rating != 3); else stage 3 (all rows).
Filtering runs on whatever device the batch lives on (index ops only); row ORDER follows the
reference (`torch.cat([where(r < lo), where(r > hi)])`: negatives first, then positives) so the
per-sample loss vector matches row for row. `MyLoss` derives the 0/1 label from (rating, stage).
"""
from __future__ import annotations

import torch


def stage_for_epoch(epoch: int, curriculums) -> int:
    if epoch < curriculums[0]:
        return 1
    if epoch < curriculums[1]:
        return 2
    return 3


def stage_row_indices(ratings: torch.Tensor, stage: int) -> torch.Tensor:
    if stage == 1:
        return torch.cat([torch.where(ratings < 2)[0], torch.where(ratings > 4)[0]])
    if stage == 2:
        return torch.cat([torch.where(ratings < 3)[0], torch.where(ratings > 3)[0]])
    return torch.arange(len(ratings), device=ratings.device)


def filter_batch(batch: dict, stage: int, device=None):
    """Returns the stage-filtered batch (moved to `device`, non-blocking) or None when no row
    survives (the reference `continue`s, src/train.py:184-185)."""
    idxs = stage_row_indices(batch["rating"], stage)
    if len(idxs) == 0:
        return None
    out = {}
    for k, v in batch.items():
        v = v[idxs]
        out[k] = v.to(device, non_blocking=True) if device is not None else v
    return out
