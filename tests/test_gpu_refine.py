"""The fused refinement chain (csrc/refine_fused.cu: all R steps of model/imf_vad.py:146-149 in one persistent kernel,
h and fp16(x) resident on chip) against the two-launch-per-step form it replaces - BIT-exact, because the multi-GPU
sharding relies on a row's result not depending on the batch it travels in and the library picks the form
by batch size - plus the range guard of the 16-bit plans and the explicit weight refresh.  Needs a B200."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import iefvad_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    assert torch.cuda.is_available()
    import iefvad_b200
    from iefvad_b200 import synth
    m = synth.build_model(iefvad_b200.MMFMIL, seed=0).eval()
    synth.perturb_(m, seed=1, scale=0.1)                 # non-zero biases everywhere (b1 / b2 of every step matter)
    return m.cuda(), synth


def _forward(m, img, ev, fused):
    m.temporal.refine_fused = fused
    with torch.no_grad():
        out = m(img, ev, None, None, None)
    return {k: v.clone() for k, v in out.items()}


@pytest.mark.parametrize("shape", [(1, 40), (3, 256), (20, 256), (150, 256), (1, 5000)])
@pytest.mark.parametrize("mode", [True])
def test_fused_chain_bit_exact_against_per_step_gemms(env, shape, mode):
    """Shapes cover one partial tile, ragged last tiles, fewer tiles than CTA pairs (every step of a tile on the same
    pair), several tiles per pair (150 x 256 = 38 400 rows = 150 tiles on 74 pairs) and a direct long call."""
    m, synth = env
    m.temporal.precision = "HH"
    B, T = shape
    vids = [synth.make_video(50 + i, T) for i in range(B)]
    img = torch.stack([v[0] for v in vids]).cuda()
    ev = torch.stack([v[1] for v in vids]).cuda()
    ref = _forward(m, img, ev, False)
    got = _forward(m, img, ev, mode)
    m.temporal.refine_fused = None
    assert torch.isfinite(got["logits"]).all()
    assert torch.equal(got["fused"], ref["fused"]), float((got["fused"] - ref["fused"]).abs().max())
    assert torch.equal(got["logits"], ref["logits"])


def test_fused_chain_matches_reference_golden_c4(env):
    """Against the unmodified reference's logits on config 4 (64 zero-padded clips), fused chain forced on."""
    import iefvad_b200
    _, synth = env
    z = load_golden("full_default.npz")
    m = synth.build_model(iefvad_b200.MMFMIL, seed=0).eval().cuda()
    m.temporal.precision = "HH"
    m.temporal.refine_fused = True
    img, ev, _, _ = synth.make_c4_batch(64)
    with torch.no_grad():
        out = m(img.cuda(), ev.cuda(), None, None, None)
    assert O.score_rel_err(out["logits"].cpu().numpy().reshape(64, 256), z["c4:logits"]) < 1e-3


def test_auto_mode_picks_the_same_bits_on_the_evaluation_path(env):
    """Valid-rows evaluation forward (what bench.py times) with the chain on auto vs forced off."""
    m, synth = env
    from iefvad_b200.evaluate import Evaluator
    m.temporal.precision = "HH"
    T = synth.config_lengths("ucf")[:120]
    classes = synth.config_classes("ucf", len(T))
    ev = Evaluator(m, T, classes, synth.make_gt(T, classes), device=torch.device("cuda:0"))
    fi, fe = zip(*[synth.make_video(v, int(T[v])) for v in range(len(T))])
    ev.set_device_features(ev.chunk_features(fi), ev.chunk_features(fe))
    with torch.no_grad():
        m.temporal.refine_fused = False
        a = ev.step()["scores"].clone()
        m.temporal.refine_fused = True
        b = ev.step()["scores"].clone()
        m.temporal.refine_fused = None
        c = ev.step()["scores"].clone()
    assert torch.equal(a, b) and torch.equal(a, c)


def test_fp16_overflow_is_loud_and_falls_back(env):
    """Scaling a refinement weight until relu(W1 x + b1) leaves the fp16 range must not produce silent garbage."""
    import iefvad_b200
    _, synth = env
    from iefvad_b200.evaluate import Evaluator
    m = synth.build_model(iefvad_b200.MMFMIL, seed=0).eval().cuda()
    m.temporal.precision = "HH"
    with torch.no_grad():
        m.temporal.refinement_blocks[3][0].weight.mul_(1.0e6)
    T = np.array([300, 17, 256])
    classes = ["Abuse", "Normal", "Arson"]
    ev = Evaluator(m, T, classes, synth.make_gt(T, classes), device=torch.device("cuda:0"))
    fi, fe = zip(*[synth.make_video(v, int(T[v])) for v in range(3)])
    ev.set_device_features(ev.chunk_features(fi), ev.chunk_features(fe))
    with torch.no_grad():
        with pytest.raises(FloatingPointError):
            ev.step()
        # direct module call: the flag is there for the caller (and raises by itself with check_every_forward)
        img, evv = synth.make_video(0, 64)
        m(img[None].cuda(), evv[None].cuda(), None, None, None)
        assert not m.temporal.check_finite()
        assert m.temporal.check_finite()                     # the check clears the flag
        m.temporal.check_every_forward = True
        with pytest.raises(FloatingPointError):
            m(img[None].cuda(), evv[None].cuda(), None, None, None)
        m.temporal.check_every_forward = False
        # automatic fall-back to the bf16 plan B (fp32's exponent range): finite scores, same as running B directly
        m.temporal.on_overflow = "fallback"
        res = ev.step()
        assert ev.fell_back and torch.isfinite(res["scores"]).all()
        m.temporal.on_overflow = "raise"
        m.temporal.precision = "B"
        direct = ev.step()
        assert torch.equal(direct["scores"], res["scores"])


def test_refresh_weights_after_a_write_through_data(env):
    import iefvad_b200
    _, synth = env
    m = synth.build_model(iefvad_b200.MMFMIL, seed=0).eval().cuda()
    img, ev = synth.make_video(0, 64)
    img, ev = img[None].cuda(), ev[None].cuda()
    with torch.no_grad():
        a = m(img, ev, None, None, None)["logits"].clone()
        m.temporal.classifier.bias.data.add_(1.0)            # bypasses the version counter
        m.temporal.refresh_weights()
        b = m(img, ev, None, None, None)["logits"].clone()
        m.temporal.classifier.bias.add_(1.0)                 # in-place op on the parameter itself: seen automatically
        c = m(img, ev, None, None, None)["logits"].clone()
    assert torch.allclose(b, a + 1.0, atol=1e-5) and torch.allclose(c, a + 2.0, atol=1e-5)


def test_pad_dedup_refuses_a_row_map_that_is_not_the_prefix_map(env):
    """with pad de-duplication the rows past len are assumed to be the zero pads - a row map that selects other
    rows must be reported, not silently computed on the wrong rows."""
    m, synth = env
    m.temporal.precision = "HH"
    img, ev = synth.make_video(3, 200)
    ci, ce = synth.chunk_video(img).cuda(), synth.chunk_video(ev).cuda()
    good = torch.arange(200, dtype=torch.int32, device="cuda")
    with torch.no_grad():
        m.temporal.scores(ci, ce, valid_lengths=[200], rowmap=good)
        assert m.temporal.check_finite()
        bad = good.clone()
        bad[100:] += 20                                   # valid rows that are not a prefix of the chunk
        m.temporal.scores(ci, ce, valid_lengths=[200], rowmap=bad)
        with pytest.raises(RuntimeError, match="prefix"):
            m.temporal.check_finite()
        m.temporal.pad_dedup = False                      # the general path takes any ascending row map
        out = m.temporal.scores(ci, ce, valid_lengths=[200], rowmap=bad)
        m.temporal.pad_dedup = True
        assert m.temporal.check_finite() and torch.isfinite(out["scores"]).all()
