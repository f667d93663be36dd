"""Drop-in for the reference's `train.loss.CLAS2` (train/loss.py:18-30) on the GPU: sigmoid, per-row top-k
(k = int(len/16 + 1)) mean and BCE in two kernels, k taken from the device `lengths` (no per-row host sync,
no O(B^2) torch.cat)."""
from __future__ import annotations

import torch

from . import ops


def CLAS2(logits: torch.Tensor, labels: torch.Tensor, lengths: torch.Tensor, device=None) -> torch.Tensor:
    """Same arguments as the reference; `device` is accepted and ignored (the tensors' device is used).
    Forward value only (this path is inference / forward-only)."""
    loss, _ = ops.clas2(logits, labels, lengths)
    return loss
