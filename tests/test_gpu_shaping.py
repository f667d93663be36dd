"""Row N2: device-side process_split / process_feat / uniform_extract vs the oracle's restatement of data/tools.py
(itself checked against the reference in the CPU suite)."""
import numpy as np
import pytest
import torch

from oracle import iefvad_oracle as O

pytestmark = pytest.mark.gpu
LENS = [1, 15, 16, 255, 256, 257, 512, 700, 1000, 4096, 3, 300]


@pytest.fixture(scope="module")
def tools():
    from iefvad_b200 import tools
    return tools


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
def test_process_split_batch_equals_reference_rule(tools, dtype):
    rng = np.random.default_rng(0)
    feats = [torch.from_numpy(rng.standard_normal((t, 768)).astype(np.float32)).to(dtype) for t in LENS]
    feats[3][5, 7] = float("nan")
    feats[3][6, 8] = float("inf")
    feats[3][7, 9] = float("-inf")
    out, chunk_off, lens = tools.process_split_batch([f.cuda() for f in feats], 256, nan_to_num=True)
    out = out.cpu()
    assert list(lens) == LENS
    for v, f in enumerate(feats):
        fn = torch.nan_to_num(f, nan=0.0)                                   # train/ucf_test.py:83-88
        ref, n = O.process_split(fn.float().numpy(), 256)
        ref = ref.reshape(-1, 256, 768)
        got = out[chunk_off[v]:chunk_off[v + 1]].float().numpy()
        assert n == LENS[v] and got.shape == ref.shape                      # incl. the extra zero chunk at T % 256 == 0
        assert np.array_equal(got, ref)
    single, n = tools.process_split(feats[1].cuda(), 256)
    assert single.shape == (256, 768) and n == 15                           # the reference returns 2-D when T < length


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_process_feat_and_uniform_extract(tools, dtype):
    rng = np.random.default_rng(1)
    lens = [1, 100, 256, 257, 300, 511, 512, 513, 1000, 4096, 10001]
    feats = [torch.from_numpy(rng.standard_normal((t, 128)).astype(np.float32)).to(dtype) for t in lens]
    out, out_len = tools.process_feat_batch([f.cuda() for f in feats], 256)
    out, out_len = out.cpu().numpy(), out_len.cpu().numpy()
    for v, f in enumerate(feats):
        ref, n = O.process_feat(f.numpy(), 256)
        assert int(out_len[v]) == n
        ref = np.asarray(ref, dtype=np.float32)
        # bins of the int32 linspace edges; the mean is a float32 sum in row order divided by the count
        np.testing.assert_allclose(out[v], ref, rtol=2e-6 if dtype == torch.float32 else 1e-3, atol=1e-7)
    ue = tools.uniform_extract(feats[4].cuda(), 256).cpu().numpy()
    np.testing.assert_allclose(ue, np.asarray(O.uniform_extract(feats[4].numpy(), 256), dtype=np.float32),
                               rtol=2e-6 if dtype == torch.float32 else 1e-3, atol=1e-7)


def test_linspace_edges_match_numpy_for_many_lengths(tools):
    """The empty-bin / bin-boundary logic depends on np.linspace(0, T, 257, dtype=int32) exactly: feed one-hot rows so
    that any off-by-one edge changes the output."""
    for T in (257, 258, 300, 383, 511, 513, 767, 1023, 1025, 5000, 65537):
        f = torch.arange(T, dtype=torch.float32)[:, None].repeat(1, 8)
        got = tools.uniform_extract(f.cuda(), 256).cpu().numpy()[:, 0]
        r = np.linspace(0, T, 257, dtype=np.int32)
        ref = np.array([f[r[i]:r[i + 1], 0].mean() if r[i] != r[i + 1] else f[r[i], 0] for i in range(256)], dtype=np.float32)
        np.testing.assert_allclose(got, ref, rtol=1e-6)
