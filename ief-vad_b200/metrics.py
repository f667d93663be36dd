"""Frame-level AUC / AP on the GPU - what the reference computes with scikit-learn on 16x-repeated segment
scores (train/ucf_test.py:151-152, :164-178, :336-353)."""
from __future__ import annotations

from typing import Tuple

import torch

from . import ops


def segment_positives(gt, repeat: int = 16, device=None) -> torch.Tensor:
    """gt: 0/1 per raw frame (numpy or tensor), `repeat` frames per embedding row -> int32 positives per row."""
    g = torch.as_tensor(gt)
    if device is not None:
        g = g.to(device)
    if g.numel() % repeat:
        raise ValueError(f"gt length {g.numel()} is not a multiple of repeat={repeat}")
    return g.reshape(-1, repeat).to(torch.int32).sum(dim=1, dtype=torch.int32)


def frame_auc_ap(scores: torch.Tensor, pos: torch.Tensor, repeat: int = 16) -> Tuple[float, float]:
    """(roc_auc_score, average_precision_score) of np.repeat(scores, repeat) vs the frame labels summarised by
    `pos` (see segment_positives).  One device->host read of 4 doubles."""
    out = ops.auc_ap(scores, pos, repeat).cpu()
    return float(out[0]), float(out[1])
