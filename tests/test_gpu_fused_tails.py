"""The fused encoder tail (csrc/outproj_ln.cu) and the fused heads + fusion kernel (csrc/heads_fuse.cu) inside the model:
switching either off (module attributes `outproj_ln`, `heads_fuse`) must not change what the evaluation forward returns -
bit for bit for heads_fuse (same arithmetic, instruction for instruction), to fp32 LayerNorm-statistics rounding for
outproj_ln.  Needs a B200."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _copy(d):
    return {k: v.clone() for k, v in d.items() if torch.is_tensor(v)}


@pytest.fixture(scope="module")
def setup():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import iefvad_b200
    from iefvad_b200 import synth
    m = synth.build_model(iefvad_b200.MMFMIL, seed=0).eval()
    synth.perturb_(m, seed=1, scale=0.1)          # non-trivial biases and LayerNorm affines
    m = m.cuda()
    vids = [synth.make_video(70 + i, T) for i, T in enumerate((300, 256, 1, 700, 255, 4096, 33))]
    ci = torch.cat([synth.chunk_video(v[0]) for v in vids]).cuda()
    ce = torch.cat([synth.chunk_video(v[1]) for v in vids]).cuda()
    valid = []
    for v in vids:
        T = v[0].shape[0]
        valid += [256] * (T // 256) + [T % 256]          # process_split (data/tools.py:100-114): T // 256 + 1 chunks
    assert ci.shape[0] == len(valid)
    rowmap = torch.cat([torch.arange(n) + c * 256 for c, n in enumerate(valid)]).to(torch.int32).cuda()
    return m, ci, ce, valid, rowmap


@pytest.mark.parametrize("dedup", [True, False])
def test_heads_fuse_is_bit_identical_to_heads_gemms_plus_fusion_kernel(setup, dedup):
    m, ci, ce, valid, rowmap = setup
    t = m.temporal
    t.pad_dedup = dedup
    with torch.no_grad():
        t.heads_fuse = True
        a = _copy(t.scores(ci, ce, None, valid, rowmap))
        t.heads_fuse = False
        b = t.scores(ci, ce, None, valid, rowmap)
        t.heads_fuse = True
    torch.cuda.synchronize()
    assert torch.isfinite(a["scores"]).all()
    assert torch.equal(a["scores"], b["scores"]) and torch.equal(a["logits"], b["logits"])


def test_extras_switch_the_fused_heads_off_and_keep_the_scores(setup):
    """Asking for mu / w means needs the tensors heads_fuse never writes: the library takes the three-launch form, and the
    scores stay the same bits."""
    m, ci, ce, valid, rowmap = setup
    t = m.temporal
    t.pad_dedup = True
    n = int(sum(valid))
    with torch.no_grad():
        a = _copy(t.scores(ci, ce, None, valid, rowmap))
        extra = {"wi_mean": torch.empty(n, device="cuda"), "we_mean": torch.empty(n, device="cuda"),
                 "image_mu": torch.empty(n, 768, device="cuda")}
        b = t.scores(ci, ce, None, valid, rowmap, extra=extra)
    torch.cuda.synchronize()
    assert torch.equal(a["scores"], b["scores"])
    assert torch.isfinite(extra["image_mu"]).all() and (extra["wi_mean"] + extra["we_mean"] - 1).abs().max() < 1e-4


@pytest.mark.parametrize("dedup", [True, False])
def test_outproj_ln_matches_gemm_plus_layernorm_launches(setup, dedup):
    m, ci, ce, valid, rowmap = setup
    t = m.temporal
    t.pad_dedup = dedup
    with torch.no_grad():
        t.outproj_ln = True
        a = _copy(t.scores(ci, ce, None, valid, rowmap))
        full_a = _copy(t(ci[:5], ce[:5], with_scores=True))
        t.outproj_ln = False
        b = t.scores(ci, ce, None, valid, rowmap)
        full_b = t(ci[:5], ce[:5], with_scores=True)
        t.outproj_ln = True
    torch.cuda.synchronize()
    # same fp16 operands and fp32 accumulation; the LayerNorm statistics are summed in a different order and the residual
    # stream between the layers is an fp16 pair (22 bits) instead of fp32: a few 1e-6 on the normalised rows, which the fp16
    # operand rounding of the next GEMM turns into one fp16 ulp on isolated elements - the two forms differ by a fraction of
    # the plan's own distance from the fp32 reference (2.2e-4, tests/test_full_configs.py)
    rel = ((a["scores"] - b["scores"]).abs() / b["scores"]).max().item()
    assert rel < 2e-4, rel
    for k in ("fused", "image_mu", "event_logvar", "w_i"):
        err = (full_a[k] - full_b[k]).abs().max().item() / full_b[k].abs().max().item()
        assert err < 5e-4, (k, err)
