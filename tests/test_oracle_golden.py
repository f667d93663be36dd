"""Pin the CPU oracle (oracle/iefvad_oracle.py) against outputs of the unmodified reference that
tests/golden/make_golden.py recorded in the authoring container.  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, params_from_npz
from oracle import iefvad_oracle as O

OUT_KEYS = ["fused", "logits", "image_mu", "event_mu", "image_logvar", "event_logvar", "w_i", "w_e"]


@pytest.mark.parametrize("name", ["small_studentt", "small_gaussian", "small_r0"])
def test_forward_small_models_all_tensors(name):
    z = load_golden(name + ".npz")
    P = params_from_npz(z)
    kw = dict(heads=int(z["heads"]), lambda_ref=float(z["lambda_ref"]), noise_model=str(z["noise_model"]),
              nu=float(z["nu"]))
    for dtype, tol in ((np.float32, 2e-5), (np.float64, 2e-6)):
        out = O.forward(P, z["img"], z["ev"], dtype=dtype, **kw)
        for k in OUT_KEYS:
            ref = z["out:" + k]
            assert out[k].shape == ref.shape
            assert O.max_norm_err(out[k], ref) < tol, (k, dtype)
        assert O.score_rel_err(out["logits"], z["out:logits"]) < tol * 5


def test_forward_unknown_noise_model_raises():
    z = load_golden("small_r0.npz")
    with pytest.raises(ValueError, match="Unsupported noise_model"):
        O.forward(params_from_npz(z), z["img"], z["ev"], heads=4, noise_model="Laplace")


def test_clas2_matches_reference():
    z = load_golden("clas2.npz")
    loss, v = O.clas2(z["logits"], z["labels"], z["lengths"])
    assert abs(float(loss) - float(z["loss"])) < 1e-6 * max(1.0, abs(float(z["loss"])))
    loss64, _ = O.clas2(z["logits"], z["labels"], z["lengths"], dtype=np.float64)
    assert abs(float(loss64) - float(z["loss"])) < 1e-6


def test_auc_ap_restatement_matches_sklearn_golden():
    z = load_golden("sklearn_auc.npz")
    for i in range(int(z["n"])):
        s, pos = z[f"{i}:scores"], z[f"{i}:pos"]
        gt = (np.arange(16)[None, :] < pos[:, None]).astype(np.float64).reshape(-1)
        rep = np.repeat(s.astype(np.float64), 16)
        auc, ap = O.roc_auc_score(gt, rep), O.average_precision_score(gt, rep)
        auc2, ap2 = O.auc_ap_segments(s, pos, 16)
        for got in (auc, auc2):
            if np.isnan(z[f"{i}:auc"]):
                assert np.isnan(got)
            else:
                assert abs(got - float(z[f"{i}:auc"])) < 1e-12
        for got in (ap, ap2):
            assert abs(got - float(z[f"{i}:ap"])) < 1e-12


def test_auc_ap_against_live_sklearn_with_ties():
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(0)
    for n in (1, 2, 17, 400):
        s = np.round(rng.random(n), 1).astype(np.float32)
        pos = rng.integers(0, 17, n)
        gt = (np.arange(16)[None, :] < pos[:, None]).astype(np.float64).reshape(-1)
        if gt.min() == gt.max():
            continue
        rep = np.repeat(s.astype(np.float64), 16)
        auc, ap = O.auc_ap_segments(s, pos, 16)
        assert abs(auc - sk.roc_auc_score(gt, rep)) < 1e-12
        assert abs(ap - sk.average_precision_score(gt, rep)) < 1e-12


def test_layers_match_reference():
    z = load_golden("layers.npz")
    x = z["sim:x"]
    assert O.max_norm_err(O.similarity_adj(x, z["sim:w0"], None), z["sim:out_none"]) < 1e-5
    assert O.max_norm_err(O.similarity_adj(x, z["sim:w0"], z["sim:seq_len"]), z["sim:out_len"]) < 1e-5
    out = O.graph_convolution(x, z["gc:adj"], z["gc:w"], z["gc:b"], residual="identity")
    assert O.max_norm_err(out, z["gc:out"]) < 1e-5
    out = O.graph_convolution(x, z["gc:adj"], z["gc2:w"], None, residual="conv", conv_w=z["gc2:conv_w"],
                              conv_b=z["gc2:conv_b"])
    assert O.max_norm_err(out, z["gc2:out"]) < 1e-5


def test_distance_adj_scan_equals_dense():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 64, 16)).astype(np.float64)
    w = rng.standard_normal((16, 8)).astype(np.float64)
    adj = O.distance_adj(2, 64, dtype=np.float64)
    dense = O.graph_convolution(x, adj, w, None, residual="none")
    scan = O.graph_convolution_distance_scan(x, w)
    assert O.max_norm_err(scan, dense) < 1e-6      # adjacency itself is built in fp32 like the reference


def test_transformer_matches_reference():
    z = load_golden("layers.npz")
    P = params_from_npz(z, "tr:param:")
    blocks = []
    i = 0
    while f"resblocks.{i}.ln_1.weight" in P:
        pre = f"resblocks.{i}."
        blocks.append({k[len(pre):]: v for k, v in P.items() if k.startswith(pre)})
        i += 1
    heads = int(z["tr:heads"])
    out = O.transformer(z["tr:x"], blocks, heads, padding_mask=None, attn_mask=z["tr:mask"])
    assert O.max_norm_err(out, z["tr:out_nopad"]) < 2e-5
    out = O.transformer(z["tr:x"], blocks, heads, padding_mask=z["tr:pad"], attn_mask=z["tr:mask"])
    assert O.max_norm_err(out, z["tr:out_masked"]) < 2e-5


def test_process_split_rules():
    for T, S in ((1, None), (255, None), (256, 2), (257, 2), (512, 3), (700, 3)):
        f = np.arange(T * 2, dtype=np.float32).reshape(T, 2) + 1
        out, n = O.process_split(f, 256)
        assert n == T
        if S is None:
            assert out.shape == (256, 2) and np.all(out[T:] == 0) and np.all(out[:T] == f)
        else:
            assert out.shape == (S, 256, 2)
            flat = out.reshape(-1, 2)
            assert np.all(flat[:T] == f) and np.all(flat[T:] == 0)


def test_input_shaping_restatement_matches_reference_tools():
    """oracle process_split / process_feat / uniform_extract == data/tools.py:65-114 run as-is (tests/golden/tools.npz)."""
    z = load_golden("tools.npz")
    for dt_name, dt in (("f32", np.float32), ("f16", np.float16)):
        rng = np.random.default_rng(12)
        for t in z["lens"]:
            t = int(t)
            feat = rng.standard_normal((t, 16)).astype(dt)
            sp, n = O.process_split(feat, 256)
            assert n == int(z[f"{dt_name}:{t}:split_len"]) and sp.shape == z[f"{dt_name}:{t}:split"].shape
            assert np.array_equal(sp, z[f"{dt_name}:{t}:split"])
            pf, m = O.process_feat(feat, 256)
            assert m == int(z[f"{dt_name}:{t}:feat_len"])
            assert np.array_equal(np.asarray(pf, dtype=np.float32), np.asarray(z[f"{dt_name}:{t}:feat"], dtype=np.float32))


@pytest.mark.parametrize("name", ["rand", "smooth", "static", "ties"])
def test_event_synthesis_restatement_matches_reference(name):
    """Row N4: extracting/ucf_gen_event.py:21-37,91-95 run as-is (golden) vs the oracle - counts bit-exact, including
    the pixel pairs whose gray difference sits on a threshold (decided by the fp32 evaluation order)."""
    import hashlib
    from iefvad_b200.synth import make_event_frames
    z = load_golden("event.npz")
    frames = make_event_frames(name)
    assert hashlib.sha256(frames.tobytes()).hexdigest() == str(z[f"{name}:sha256"])
    for thr, clamp in ((25, 10), (10, 10), (25, 3)):
        assert np.array_equal(O.generate_event_image(frames, thr), z[f"{name}:{thr}:sum"].astype(np.float32))
        ev = O.event_images(frames, thr, clamp)
        assert ev.shape == (frames.shape[0], 3) + frames.shape[2:4]
        for c in range(3):
            assert np.array_equal(ev[:, c], z[f"{name}:{thr}:{clamp}:event"], equal_nan=True)


def test_localisation_map_oracle_matches_reference_metrics():
    """Row N5: oracle.loc_map (restating train/metrics.py:44-126) against the reference's own getDetectionMAP output."""
    from iefvad_b200 import synth
    z = load_golden("locmap.npz")
    preds, segs, labels = synth.make_locmap_case()
    for th, want in zip(z["ious"], z["dmap"]):
        assert abs(O.loc_map(preds, float(th), segs, labels) - float(want)) < 1e-9
    assert abs(O.loc_map(preds[:9], 0.3, segs[:9], labels[:9]) - float(z["short"])) < 1e-9
    dead = [p.copy() for p in preds]
    for p in dead:
        p[:, 5] = -1.0                                  # no proposal for class 5 anywhere -> the reference returns 0 (:92-93)
    assert O.loc_map(dead, 0.3, segs, labels) == 0
