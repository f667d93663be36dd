"""Row N4 on the GPU: event-frame synthesis (extracting/ucf_gen_event.py:21-37,91-95) through the C ABI
(`iefvad_event_image`) against the reference's golden outputs and the oracle - integer counts, so bit-exact."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import iefvad_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ext():
    from iefvad_b200 import extracting, synth
    return extracting, synth


@pytest.mark.parametrize("name", ["rand", "smooth", "static", "ties"])
def test_event_images_match_reference_golden(ext, name):
    extracting, synth = ext
    z = load_golden("event.npz")
    frames = synth.make_event_frames(name)
    assert hashlib.sha256(frames.tobytes()).hexdigest() == str(z[f"{name}:sha256"])
    dev_frames = torch.from_numpy(frames).cuda()
    for thr, clamp in ((25, 10), (10, 10), (25, 3)):
        s = extracting.generate_event_image(dev_frames, thr).cpu().numpy()
        assert s.dtype == np.float32 and np.array_equal(s, z[f"{name}:{thr}:sum"].astype(np.float32))
        ev = extracting.event_images(dev_frames, thr, clamp).cpu().numpy()
        assert ev.shape == (frames.shape[0], 3) + frames.shape[2:4]
        for c in range(3):
            assert np.array_equal(ev[:, c], z[f"{name}:{thr}:{clamp}:event"], equal_nan=True), (thr, clamp, c)


@pytest.mark.parametrize("shape", [(1, 2, 1, 1), (2, 16, 5, 7), (3, 5, 8, 12), (1, 16, 224, 224), (0, 16, 4, 4)])
def test_event_images_match_oracle_on_odd_shapes(ext, shape):
    """Ragged sizes (H*W not a multiple of 4 takes the scalar path), a single difference, an empty batch, host input."""
    extracting, _ = ext
    B, C, H, W = shape
    rng = np.random.default_rng(B * 1000 + H * W)
    base = rng.integers(0, 256, (B, 1, H, W, 3))
    frames = np.clip(base + rng.normal(0, 12, (B, C, H, W, 3)), 0, 255).astype(np.uint8)
    got = extracting.generate_event_image(frames, 10, device="cuda").cpu().numpy()
    assert np.array_equal(got, O.generate_event_image(frames, 10).reshape(B, H, W))
    if B:
        ev = extracting.event_images(torch.from_numpy(frames).cuda(), 10, 4).cpu().numpy()
        assert np.array_equal(ev, O.event_images(frames, 10, 4), equal_nan=True)


def test_event_rejects_bad_input(ext):
    extracting, _ = ext
    with pytest.raises(RuntimeError, match="uint8"):
        extracting.generate_event_image(torch.zeros(1, 2, 4, 4, 3).cuda())
    with pytest.raises(RuntimeError, match=r"\[B, C, H, W, 3\]"):
        extracting.generate_event_image(torch.zeros(2, 4, 4, 3, dtype=torch.uint8).cuda())
    with pytest.raises(RuntimeError, match="no CPU path"):
        extracting.generate_event_image(np.zeros((1, 2, 4, 4, 3), dtype=np.uint8))


def test_event_throughput_is_hbm_bound(ext):
    """One reference batch of 32 stacks (extracting/ucf_gen_event.py batch_size x chunk_size x 224 x 224 x 3): the kernels
    move 3 C + 16 bytes per pixel; report-level check only that it runs at a sane fraction of HBM."""
    extracting, _ = ext
    frames = torch.randint(0, 256, (32, 16, 224, 224, 3), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        extracting.event_images(frames)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        extracting.event_images(frames)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = 32 * 224 * 224 * (3 * 16 + 4 + 4 + 12) / ms / 1e6
    print(f"event_images 32 x 16 x 224 x 224: {ms * 1e3:.1f} us, {gbs:.0f} GB/s")
    assert gbs > 500
