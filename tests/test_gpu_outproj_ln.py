"""The fused out-projection + residual + LayerNorm(s) kernel (csrc/outproj_ln.cu) against a float64 restatement of
model/imf_vad.py:115-117 on the same fp16-rounded operands.  Needs a B200."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

D = 768


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from iefvad_b200 import ops as _ops
    return _ops


def _ln(x, w, b, eps=1e-5):
    m = x.mean(-1, keepdim=True)
    v = ((x - m) ** 2).mean(-1, keepdim=True)
    return (x - m) / torch.sqrt(v + eps) * w + b


def _case(rows, seed, double, offset=0.0):
    g = torch.Generator().manual_seed(seed)
    ctx = torch.randn(rows, D, generator=g)
    w = torch.randn(D, D, generator=g) * D ** -0.5
    bias = torch.randn(D, generator=g) * 0.1
    resid = torch.randn(rows, D, generator=g) * 1.5 + offset
    lw = [1 + 0.2 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)]
    lw2 = [1 + 0.2 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)] if double else [None, None]
    y = resid.double() + ctx.half().double() @ w.half().double().T + bias.double()
    ref = _ln(y, lw[0].double(), lw[1].double())
    if double:
        ref = _ln(ref, lw2[0].double(), lw2[1].double())
    return ctx, w, bias, resid, lw, lw2, ref


@pytest.mark.parametrize("rows", [1, 200, 256, 777, 5000, 40000])
@pytest.mark.parametrize("double", [False, True])
def test_outproj_ln_matches_fp64(ops, rows, double):
    ctx, w, bias, resid, lw, lw2, ref = _case(rows, rows + int(double), double, offset=0.7)
    hi, lo = ops.outproj_ln(ctx.cuda(), w.cuda(), bias.cuda(), resid.cuda(), lw[0].cuda(), lw[1].cuda(),
                            None if lw2[0] is None else lw2[0].cuda(), None if lw2[1] is None else lw2[1].cuda())
    torch.cuda.synchronize()
    got = hi.double().cpu() + lo.double().cpu()
    scale = ref.abs().max().item()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() < 2e-5 * scale          # fp32 accumulation / statistics, 22-bit output pair
    # hi alone is the fp16 rounding of the result
    assert (hi.double().cpu() - ref).abs().max().item() < 6e-4 * scale


def test_outproj_ln_row_map_and_batch_invariance(ops):
    rows = 3000
    ctx, w, bias, resid, lw, lw2, ref = _case(rows, 5, True)
    args = [t.cuda() for t in (ctx, w, bias, resid, lw[0], lw[1], lw2[0], lw2[1])]
    hi, _ = ops.outproj_ln(*args, want_lo=False)
    # keep every third row, compacted
    keep = torch.arange(rows) % 3 == 0
    rmap = torch.full((rows,), -1, dtype=torch.int32)
    rmap[keep] = torch.arange(int(keep.sum()), dtype=torch.int32)
    hm, lo = ops.outproj_ln(*args, row_map=rmap.cuda(), out_rows=int(keep.sum()))
    assert lo is None
    assert torch.equal(hm.cpu(), hi.cpu()[keep])
    # a row's result does not depend on where it sits in the batch nor on the batch size (multi-GPU determinism)
    sl = slice(517, 517 + 300)
    h2, _ = ops.outproj_ln(args[0][sl].contiguous(), args[1], args[2], args[3][sl].contiguous(), *args[4:], want_lo=False)
    assert torch.equal(h2.cpu(), hi.cpu()[sl])
