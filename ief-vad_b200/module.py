"""B200 counterparts of the reference's `model/module.py` (CLIP-style temporal transformer): same class names,
constructor arguments and state_dict layout (`resblocks.{i}.{ln_1,ln_2}.{weight,bias}`, `attn.in_proj_*`,
`attn.out_proj.*`, `mlp.c_fc.*`, `mlp.c_proj.*`), same tuple-in / tuple-out forward on SEQ-FIRST tensors.
The whole stack runs in one libiefvad.so call (`iefvad_transformer`).  Inference only."""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib
from .ops import PLAN_CODES, _f32c, _stream


class LayerNorm(nn.LayerNorm):
    """module.py:7-12: LayerNorm computed in fp32 whatever the input dtype.  Inside `Transformer` the arithmetic happens
    in the fused call; called on its own it runs the library's LayerNorm kernel."""

    def forward(self, x: torch.Tensor):
        from . import ops
        if tuple(self.normalized_shape) != (x.shape[-1],):
            raise RuntimeError("LayerNorm over the last dimension only")
        out = ops.layernorm(x, self.weight, self.bias, eps=self.eps)           # upcasts to fp32 (:9-11)
        return out.to(x.dtype)


class QuickGELU(nn.Module):
    """module.py:15-17: x * sigmoid(1.702 x).  Inside `Transformer` it is the c_fc GEMM's epilogue; called on its own it
    runs the library's element-wise kernel."""

    def forward(self, x: torch.Tensor):
        from . import ops
        return ops.quickgelu(x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask
        self.n_head = n_head

    def _params(self):
        return [self.ln_1.weight, self.ln_1.bias, self.attn.in_proj_weight, self.attn.in_proj_bias,
                self.attn.out_proj.weight, self.attn.out_proj.bias, self.ln_2.weight, self.ln_2.bias,
                self.mlp.c_fc.weight, self.mlp.c_fc.bias, self.mlp.c_proj.weight, self.mlp.c_proj.bias]

    def forward(self, x):
        return _run([self], x)


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])
        self.precision = "split"

    def forward(self, x):
        return _run(list(self.resblocks), x, getattr(self, "precision", "split"))


def _run(blocks, x, precision: str = "split"):
    x, padding_mask = x                                                    # module.py:40
    xs = _f32c(x, "Transformer")
    L, N, D = xs.shape
    mask = blocks[0].attn_mask
    if any((b.attn_mask is None) != (mask is None) or (mask is not None and b.attn_mask is not mask
                                                      and not torch.equal(b.attn_mask, mask)) for b in blocks):
        raise RuntimeError("all blocks of one call must share the attn_mask")
    m = _f32c(mask.to(xs.device), "Transformer") if mask is not None else None
    kp = padding_mask.to(device=xs.device, dtype=torch.uint8).contiguous() if padding_mask is not None else None
    keep = [p.detach().to(device=xs.device, dtype=torch.float32).contiguous() for b in blocks for p in b._params()]
    table = (C.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
    out = torch.empty_like(xs)
    with torch.cuda.device(xs.device):
        _lib.check(_lib.lib.iefvad_transformer(xs.data_ptr(), C.cast(table, C.c_void_p), len(blocks), L, N, D,
                                               blocks[0].n_head, _lib.ptr(m), _lib.ptr(kp), PLAN_CODES[precision],
                                               out.data_ptr(), _stream(xs)))
    return (out.to(x.dtype), padding_mask)
