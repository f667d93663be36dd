"""B200 counterparts of the graph classes in the reference's `model/layers.py` (same class names, constructor
arguments, parameter names / shapes / initialisation and forward signatures):

    GraphConvolution (layers.py:64-111), SimilarityAdj (:114-163), DistanceAdj (:166-179)

They are dead code in the reference's live graph but named by the north star; each forward is one
call into libiefvad.so (include/iefvad.h, `iefvad_graph_convolution`, `iefvad_similarity_adj`,
`iefvad_distance_adj`).  `precision` selects "fp32" (FFMA), "bf16" or "split" (tcgen05, default).  Inference only."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from .ops import PLAN_CODES, _f32c, _stream


def _wt(w: torch.Tensor) -> torch.Tensor:
    return w.detach().float().t().contiguous()       # the library takes [out, in]; the reference stores [in, out]


class GraphConvolution(nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = False, residual: bool = True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(in_features, out_features))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        nn.init.xavier_uniform_(self.weight)                       # layers.py:85-89
        if self.bias is not None:
            self.bias.data.fill_(0.1)
        self.residual_mode = 0 if not residual else (1 if in_features == out_features else 2)
        if self.residual_mode == 2:                                # layers.py:84
            self.residual = nn.Conv1d(in_features, out_features, kernel_size=5, padding=2)
        self.precision = "split"

    def forward(self, input: torch.Tensor, adj: Optional[torch.Tensor]) -> torch.Tensor:
        """adj [B, T, T]; adj=None evaluates the DistanceAdj adjacency as a bidirectional scan (no T x T matrix)."""
        x = _f32c(input, "GraphConvolution")
        B, T, Din = x.shape
        a = _f32c(adj, "GraphConvolution") if adj is not None else None
        if a is not None and tuple(a.shape) != (B, T, T):
            raise RuntimeError(f"adj must be [{B}, {T}, {T}], got {tuple(a.shape)}")
        out = torch.empty((B, T, self.out_features), dtype=torch.float32, device=x.device)
        wt = _wt(self.weight)
        bias = self.bias.detach().float().contiguous() if self.bias is not None else None
        cw = cb = None
        if self.residual_mode == 2:
            cw = self.residual.weight.detach().float().permute(0, 2, 1).contiguous()      # [out, 5, in]
            cb = self.residual.bias.detach().float().contiguous()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib.iefvad_graph_convolution(
                x.data_ptr(), _lib.ptr(a), wt.data_ptr(), _lib.ptr(bias), self.residual_mode, _lib.ptr(cw), _lib.ptr(cb),
                B, T, Din, self.out_features, PLAN_CODES[self.precision], out.data_ptr(), _stream(x)))
        return out

    def __repr__(self):
        return f"{self.__class__.__name__} ({self.in_features} -> {self.out_features})"


class SimilarityAdj(nn.Module):
    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight0 = nn.Parameter(torch.empty(in_features, out_features))
        self.weight1 = nn.Parameter(torch.empty(in_features, out_features))   # never read by forward (layers.py:132-133)
        self.register_parameter("bias", None)
        nn.init.xavier_uniform_(self.weight0)
        nn.init.xavier_uniform_(self.weight1)
        self.precision = "split"

    def forward(self, input: torch.Tensor, seq_len: Optional[Sequence[int]]) -> torch.Tensor:
        x = _f32c(input, "SimilarityAdj")
        B, T, Din = x.shape
        out = torch.empty((B, T, T), dtype=torch.float32, device=x.device)
        wt = _wt(self.weight0)
        lens = None
        if seq_len is not None:
            if len(seq_len) != B:
                # the reference loops `for i in range(len(seq_len))` and leaves the other batch elements zero
                raise RuntimeError(f"seq_len has {len(seq_len)} entries for a batch of {B}")
            lens = (C.c_int64 * B)(*[int(v) for v in seq_len])
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib.iefvad_similarity_adj(
                x.data_ptr(), wt.data_ptr(), C.cast(lens, C.c_void_p) if lens is not None else None, B, T, Din,
                self.out_features, PLAN_CODES[self.precision], out.data_ptr(), _stream(x)))
        return out

    def __repr__(self):
        return f"{self.__class__.__name__} ({self.in_features} -> {self.out_features})"


class DistanceAdj(nn.Module):
    def __init__(self):
        super().__init__()
        self.sigma = nn.Parameter(torch.full((1,), 0.1))           # unused by forward, as in the reference (:169-170)

    def forward(self, batch_size: int, max_seqlen: int) -> torch.Tensor:
        dev = self.sigma.device if self.sigma.is_cuda else torch.device("cuda")    # the reference hard-codes 'cuda'
        out = torch.empty((batch_size, max_seqlen, max_seqlen), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.iefvad_distance_adj(batch_size, max_seqlen, out.data_ptr(),
                                                    torch.cuda.current_stream(dev).cuda_stream))
        return out


def distance_scan(s: torch.Tensor) -> torch.Tensor:
    """DistanceAdj(B, T) @ s without the T x T matrix: s [B, T, D] -> [B, T, D]."""
    s = _f32c(s, "distance_scan")
    B, T, D = s.shape
    y = torch.empty_like(s)
    with torch.cuda.device(s.device):
        _lib.check(_lib.lib.iefvad_distance_scan(s.data_ptr(), B, T, D, y.data_ptr(), _stream(s)))
    return y
