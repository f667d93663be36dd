"""Drop-in for the reference's `train.loss.CLAS2` (train/loss.py:18-30) on the GPU: sigmoid, per-row top-k
(k = int(len/16 + 1)) mean and BCE in two kernels, k taken from the device `lengths` (no per-row host sync,
no O(B^2) torch.cat)."""
from __future__ import annotations

import torch

from . import ops


def CLAS2(logits: torch.Tensor, labels: torch.Tensor, lengths: torch.Tensor, device=None) -> torch.Tensor:
    """Same arguments as the reference; `device` is accepted and ignored (the tensors' device is used).  With a
    `logits` that requires grad the result carries a grad_fn (train.Clas2Fn: the top-k scatter of the BCE gradient),
    so `loss.backward()` of train/ucf_train.py:105 works."""
    if torch.is_grad_enabled() and logits.requires_grad:
        from .train import Clas2Fn
        return Clas2Fn.apply(logits, labels, lengths)
    loss, _ = ops.clas2(logits, labels, lengths)
    return loss
