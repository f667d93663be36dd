"""Run the fused out-projection + LayerNorm kernel alone at the benchmark shape (for ncu): python scripts/outproj_ln_probe.py [rows] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iefvad_b200 import ops  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 78336
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
D = 768
g = torch.Generator(device="cuda").manual_seed(0)
ctx = torch.randn(rows, D, device="cuda", generator=g)
w = torch.randn(D, D, device="cuda", generator=g) * D ** -0.5
resid = torch.randn(rows, D, device="cuda", generator=g)
vec = [torch.randn(D, device="cuda", generator=g) * 0.1 + (1.0 if i in (1, 3) else 0.0) for i in range(5)]
rmap = torch.arange(rows, dtype=torch.int32, device="cuda")
for _ in range(reps):
    ops.outproj_ln(ctx, w, vec[0], resid, vec[1], vec[2])                       # layer 0 form: pair out
    ops.outproj_ln(ctx, w, vec[0], resid, vec[1], vec[2], vec[3], vec[4], row_map=rmap, out_rows=rows)   # last layer form
torch.cuda.synchronize()
print("ok")
