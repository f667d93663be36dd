"""B200 counterpart of the event-frame synthesis in the reference's extraction scripts:

    generate_event_image(frames, threshold)   extracting/ucf_gen_event.py:21-37 (same name and arguments)
    event_images(frames, threshold, clamp)    the same plus the clamp / normalise / 3-channel stack of its caller (:91-95)

`frames` is what `video_crop(...).reshape(batch, chunk, 224, 224, 3)` yields there: uint8 [B, C, H, W, 3].  Both run
as one / two HBM-bound kernels of libiefvad.so (`iefvad_event_image`); there is no CPU path."""
from __future__ import annotations

import torch

from . import _lib


def _frames(frames, device) -> torch.Tensor:
    t = torch.as_tensor(frames)
    if t.dtype != torch.uint8:
        raise RuntimeError(f"frames must be uint8 (decoded video frames), got {t.dtype}")
    if t.dim() != 5 or t.shape[-1] != 3:
        raise RuntimeError(f"frames must be [B, C, H, W, 3], got {tuple(t.shape)}")
    if not t.is_cuda:
        if device is None:
            raise RuntimeError("frames on the host need an explicit CUDA `device` (there is no CPU path)")
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def _run(frames: torch.Tensor, threshold: float, clamp: float, want_sum: bool, want_event: bool):
    B, C, H, W, _ = frames.shape
    s = torch.empty((B, H, W), dtype=torch.float32, device=frames.device) if want_sum else None
    e = torch.empty((B, 3, H, W), dtype=torch.float32, device=frames.device) if want_event else None
    with torch.cuda.device(frames.device):
        _lib.check(_lib.lib.iefvad_event_image(frames.data_ptr(), B, C, H, W, float(threshold), float(clamp), _lib.ptr(s),
                                               _lib.ptr(e), torch.cuda.current_stream(frames.device).cuda_stream))
    return s, e


def generate_event_image(frames, threshold=25, device=None) -> torch.Tensor:
    """[B, C, H, W, 3] uint8 -> [B, H, W] fp32: how many of the C - 1 gray-level frame differences exceed `threshold`."""
    return _run(_frames(frames, device), threshold, float("inf"), True, False)[0]


def event_images(frames, threshold=25, clamp=10, device=None) -> torch.Tensor:
    """[B, C, H, W, 3] uint8 -> [B, 3, H, W] fp32 in [0, 1]: clamp(counts, 0, clamp) / max over the batch, 3 channels."""
    return _run(_frames(frames, device), threshold, clamp, False, True)[1]
