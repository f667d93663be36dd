"""Device-side counterparts of the reference's feature shaping (data/tools.py:65-114): same function names and
return conventions for a single video, plus batched forms over a whole video list (one kernel launch each).

    process_split(feat, length)  -> ([S, length, D], clip_length)     data/tools.py:100-114
    process_feat(feat, length)   -> ([length, D] fp32, length_out)    data/tools.py:89-97 (is_random=False)
    uniform_extract(feat, t_max) -> [t_max, D] fp32                   data/tools.py:65-73 (avg=True)
    pad(feat, min_len)           -> zero-padded rows                  data/tools.py:81-86

Inputs are CUDA tensors (fp16 / bf16 / fp32); there is no CPU path."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .evaluate import num_chunks

_CODES = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def _pack(feats: Sequence[torch.Tensor]):
    if not feats:
        raise ValueError("empty video list")
    dev, dt, D = feats[0].device, feats[0].dtype, feats[0].shape[1]
    if not feats[0].is_cuda:
        raise RuntimeError("CUDA tensors required (the B200 path has no CPU fallback)")
    if dt not in _CODES:
        raise RuntimeError(f"unsupported dtype {dt}")
    lens = np.array([int(f.shape[0]) for f in feats], dtype=np.int64)
    packed = torch.cat([f.reshape(-1, D) for f in feats]).contiguous() if len(feats) > 1 else feats[0].contiguous()
    row_off = torch.as_tensor(np.concatenate([[0], np.cumsum(lens)]), device=dev)
    return packed, row_off, lens, dev, dt, D


def process_split_batch(feats: Sequence[torch.Tensor], length: int = 256, nan_to_num: bool = True
                        ) -> Tuple[torch.Tensor, np.ndarray, np.ndarray]:
    """All videos of a list -> ([sum S_v, length, D] in the input dtype, chunk offsets [V + 1], clip lengths [V])."""
    packed, row_off, lens, dev, dt, D = _pack(feats)
    chunks = np.array([num_chunks(int(t), length) for t in lens], dtype=np.int64)
    chunk_off = np.concatenate([[0], np.cumsum(chunks)])
    total = int(chunk_off[-1])
    out = torch.empty((total, length, D), dtype=dt, device=dev)
    coff = torch.as_tensor(chunk_off, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.iefvad_process_split(packed.data_ptr(), _CODES[dt], row_off.data_ptr(), len(lens), D, length,
                                                 coff.data_ptr(), total, out.data_ptr(), int(nan_to_num),
                                                 torch.cuda.current_stream(dev).cuda_stream))
    return out, chunk_off, lens


def process_feat_batch(feats: Sequence[torch.Tensor], length: int = 256, nan_to_num: bool = False
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
    """All videos of a list -> ([V, length, D] fp32, lengths int64 [V] on the device)."""
    packed, row_off, lens, dev, dt, D = _pack(feats)
    out = torch.empty((len(lens), length, D), dtype=torch.float32, device=dev)
    out_len = torch.empty(len(lens), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.iefvad_process_feat(packed.data_ptr(), _CODES[dt], row_off.data_ptr(), len(lens), D, length,
                                                out.data_ptr(), out_len.data_ptr(), int(nan_to_num),
                                                torch.cuda.current_stream(dev).cuda_stream))
    return out, out_len


def process_split(feat: torch.Tensor, length: int):
    out, _, lens = process_split_batch([feat], length, nan_to_num=False)
    return (out[0] if feat.shape[0] < length else out), int(lens[0])       # the reference returns 2-D when no split


def process_feat(feat: torch.Tensor, length: int, is_random: bool = False):
    if is_random and feat.shape[0] > length:
        raise NotImplementedError("random_extract (np.random) is host-side data augmentation and is not part of this path")
    out, out_len = process_feat_batch([feat], length)
    return out[0], int(min(feat.shape[0], length))


def uniform_extract(feat: torch.Tensor, t_max: int, avg: bool = True) -> torch.Tensor:
    if not avg:
        raise NotImplementedError("only the avg=True branch is used by the reference (process_feat)")
    if feat.shape[0] <= t_max:
        raise ValueError("uniform_extract expects more rows than t_max (process_feat pads shorter clips)")
    return process_feat_batch([feat], t_max)[0][0]


def pad(feat: torch.Tensor, min_len: int) -> torch.Tensor:
    if feat.shape[0] > min_len:
        return feat
    out, _, _ = process_split_batch([feat], min_len, nan_to_num=False)
    return out[0] if feat.shape[0] < min_len else out[0]
