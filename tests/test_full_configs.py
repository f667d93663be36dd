"""BASELINE configs 2, 3 and 4 (B = 128) pinned AT THEIR STATED SIZE against the unmodified reference.

The goldens (tests/golden/config2_ucf.npz, config3_xd.npz, config4_b128.npz) hold the scores the reference itself
produced on the full workloads (make_golden.py config2 / config3 / config4: its own train/ucf_test.py:test() loop for
config 2).  CPU tests: the goldens are self-consistent (sklearn arithmetic restated by the oracle reproduces the stored
AUC / AP exactly; the oracle forward reproduces sampled videos incl. the 17-chunk T = 4096 one).  GPU tests: the DEFAULT
evaluation path of the product (plan HH, valid rows, pad de-duplication, ragged host input - what bench.py times)
reproduces every score within the 1e-3 tolerance of the bf16-class plans and the AUC / AP within the bound the score
error induces."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import iefvad_oracle as O

SCORE_TOL = 1e-3      # BASELINE north_star: per-frame scores within 1e-3 relative for the 16-bit tensor-core paths


def _rel(p, ref):
    p, ref = np.asarray(p, np.float64), np.asarray(ref, np.float64)
    return float(np.max(np.abs(p - ref) / ref))


# ----------------------------------------------------------------------------------------------------- CPU
def test_config2_golden_is_self_consistent():
    from iefvad_b200 import synth
    z = load_golden("config2_ucf.npz")
    T = synth.config_lengths("ucf")
    classes = synth.config_classes("ucf", len(T))
    assert np.array_equal(z["lengths"], T) and [str(c) for c in z["classes"]] == classes
    assert z["scores"].size == int(T.sum()) == 77788 and int(T.max()) == 4096
    gt = synth.make_gt(T, classes)
    pos = gt.reshape(-1, 16).sum(axis=1).astype(np.int64)
    auc, ap = O.auc_ap_segments(z["scores"], pos, 16)
    assert abs(auc - float(z["AUC"])) < 1e-12 and abs(ap - float(z["AP"])) < 1e-12
    assert abs(float(z["ret"][0]) - float(z["AUC"])) < 1e-12 and abs(float(z["ret"][1]) - float(z["AP"])) < 1e-12
    # class-wise and Ano-AUC through the oracle's sklearn restatement
    off = np.concatenate([[0], np.cumsum(T)])
    keys = [str(k) for k in z["class_keys"]]
    for c, key in enumerate(keys):
        idx = np.concatenate([np.arange(off[v], off[v + 1]) for v in range(len(T)) if classes[v] == key])
        if pos[idx].sum() == 0:
            assert np.isnan(z["classwise"][c]).all()
            continue
        a, p = O.auc_ap_segments(z["scores"][idx], pos[idx], 16)
        assert abs(a - z["classwise"][c, 0]) < 1e-12 and abs(p - z["classwise"][c, 1]) < 1e-12, key
    idx = np.concatenate([np.arange(off[v], off[v + 1]) for v in range(len(T)) if classes[v] != "Normal"])
    a, _ = O.auc_ap_segments(z["scores"][idx], pos[idx], 16)
    assert abs(a - float(z["ano_AUC"])) < 1e-12


def test_oracle_reproduces_config2_videos_incl_17_chunks():
    """Oracle forward on the T = 4096 video (17 chunks, the last one all zeros) and two others of the bench workload."""
    from iefvad_b200 import synth
    from iefvad_b200.imf_vad import MMFMIL
    z = load_golden("config2_ucf.npz")
    m = synth.build_model(MMFMIL, seed=0).eval()
    assert synth.state_digest(m.state_dict()) == str(z["digest"])
    P = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    T = z["lengths"]
    off = np.concatenate([[0], np.cumsum(T)])
    for v in (7, 4, 0):                                  # T = 4096 (17 chunks), 256 (extra zero chunk), 1
        img, ev = synth.make_video(v, int(T[v]))
        fi, n = O.process_split(img.numpy(), 256)
        fe, _ = O.process_split(ev.numpy(), 256)
        fi = fi if fi.ndim == 3 else fi[None]
        fe = fe if fe.ndim == 3 else fe[None]
        assert fi.shape[0] == (1 if T[v] < 256 else T[v] // 256 + 1)
        out = O.forward(P, fi, fe)
        sc = O.sigmoid(out["logits"].reshape(-1)[:n])
        ref = z["scores"][off[v]:off[v + 1]]
        assert _rel(sc, ref) < 2e-5, v
        wi = out["w_i"].reshape(-1, 768)[:n].mean(axis=-1)
        assert np.max(np.abs(wi - z["wi_mean"][off[v]:off[v + 1]])) < 2e-6
        assert O.max_norm_err(out["fused"].reshape(-1, 768)[0], z["first_rows"][v, 0]) < 2e-5


def test_config3_and_config4_goldens_self_consistent():
    from iefvad_b200 import synth
    z = load_golden("config3_xd.npz")
    T = synth.config_lengths("xd")
    assert np.array_equal(z["lengths"], T) and z["scores"].size == int(T.sum()) == 684801
    gt = synth.make_gt(T, synth.config_classes("xd", len(T)))
    auc, ap = O.auc_ap_segments(z["scores"], gt.reshape(-1, 16).sum(axis=1).astype(np.int64), 16)
    assert abs(auc - float(z["AUC"])) < 1e-12 and abs(ap - float(z["AP"])) < 1e-12
    z4 = load_golden("config4_b128.npz")
    img, ev, lengths, labels = synth.make_c4_batch(128)
    assert np.array_equal(lengths.numpy(), z4["lengths"]) and np.array_equal(labels.numpy(), z4["labels"])
    loss, _ = O.clas2(z4["logits"][..., None], z4["labels"], z4["lengths"])
    assert abs(float(loss) - float(z4["loss"])) < 1e-6


# ----------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def gpu_model():
    assert torch.cuda.is_available()
    import iefvad_b200
    from iefvad_b200 import synth
    m = synth.build_model(iefvad_b200.MMFMIL, seed=0).eval().cuda()
    return m, synth


def _evaluator(m, synth, name):
    from iefvad_b200.evaluate import Evaluator
    T = synth.config_lengths(name)
    classes = synth.config_classes(name, len(T))
    ev = Evaluator(m, T, classes, synth.make_gt(T, classes), device=torch.device("cuda"))
    fi, fe = [], []
    for v in ev.mine:
        a, b = synth.make_video(int(v), int(T[v]))
        fi.append(a)
        fe.append(b)
    return ev, fi, fe


def _auc_bound(scores, pos, err):
    """|AUC - AUC_ref| is at most the mass of (positive, negative) frame pairs whose scores are closer than the
    perturbation can bridge; a cheap upper bound: pairs within 2 * err * max(score) of each other."""
    s = np.sort(np.asarray(scores, np.float64))
    width = 2 * err * s[-1]
    close = (np.searchsorted(s, s + width, side="right") - np.searchsorted(s, s - width, side="left")).mean() / s.size
    return float(close)


@pytest.mark.gpu
@pytest.mark.parametrize("inputs", ["device_chunks", "host_ragged"])
def test_config2_default_evaluator_path_full_size(gpu_model, inputs):
    """The bench workload itself through the default path (HH, valid rows, pad de-dup), both input routes."""
    m, synth = gpu_model
    z = load_golden("config2_ucf.npz")
    m.temporal.precision = "HH"
    ev, fi, fe = _evaluator(m, synth, "ucf")
    assert ev.valid_rows_only and m.temporal.pad_dedup
    with torch.no_grad():
        if inputs == "device_chunks":
            ev.set_device_features(ev.chunk_features(fi), ev.chunk_features(fe))
            res = ev.step()
        else:
            ev.set_host_ragged(fi, fe)
            res = ev.step(host_inputs=True)
    sc = res["scores"].cpu().numpy()
    err = _rel(sc, z["scores"])
    assert err < SCORE_TOL, err
    bound = _auc_bound(z["scores"], None, err) + 1e-9
    assert abs(res["AUC"] - float(z["AUC"])) <= bound, (res["AUC"], float(z["AUC"]), bound)
    assert abs(res["AP"] - float(z["AP"])) <= 4 * bound
    assert abs(res["ano_AUC"] - float(z["ano_AUC"])) <= 2 * bound
    keys = [str(k) for k in z["class_keys"]]
    for c, key in enumerate(keys):
        if np.isnan(z["classwise"][c, 0]):
            assert key not in res["classwise"]
        else:
            assert abs(res["classwise"][key][0] - z["classwise"][c, 0]) <= 8 * bound, key
            assert abs(res["classwise"][key][1] - z["classwise"][c, 1]) <= 16 * bound, key
    # given the reference's own scores the device ranking is exact
    exact = ev.metrics(torch.from_numpy(z["scores"]).cuda())
    assert abs(exact["AUC"] - float(z["AUC"])) < 1e-12 and abs(exact["AP"] - float(z["AP"])) < 1e-12
    assert abs(exact["ano_AUC"] - float(z["ano_AUC"])) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("inputs", ["device_chunks", "host_ragged"])
def test_config2_evaluator_extras_match_the_reference_loop(gpu_model, inputs):
    """Row N1: what the reference's loop collects per class besides the scores (train/ucf_test.py:124-144) - the per-frame
    means of the fusion weights (reduced inside the fusion kernel) and the frames' fused / image_mu / event_mu rows - from
    the valid-rows evaluation forward, against the reference's own values on the full bench workload."""
    m, synth = gpu_model
    z = load_golden("config2_ucf.npz")
    m.temporal.precision = "HH"
    ev, fi, fe = _evaluator(m, synth, "ucf")
    with torch.no_grad():
        if inputs == "device_chunks":
            ev.set_device_features(ev.chunk_features(fi), ev.chunk_features(fe))
            res = ev.step(extras=("w_mean", "wide"))
            plain = ev.step()
        else:
            ev.set_host_ragged(fi, fe)
            res = ev.step(host_inputs=True, extras=("w_mean", "wide"))
            plain = ev.step(host_inputs=True)
    assert torch.equal(res["scores"], plain["scores"])           # asking for the extras does not change a bit of the scores
    assert np.max(np.abs(res["wi_mean"].cpu().numpy() - z["wi_mean"])) < 1e-4
    assert np.max(np.abs(res["we_mean"].cpu().numpy() - z["we_mean"])) < 1e-4
    T = z["lengths"]
    off = np.concatenate([[0], np.cumsum(T)])[:-1]
    rows = torch.as_tensor(off, device="cuda")
    for j, k in enumerate(("fused", "image_mu", "event_mu")):
        assert res[k].shape == (int(T.sum()), 768)
        assert O.max_norm_err(res[k][rows].cpu().numpy(), z["first_rows"][:, j]) < 1e-3, k
    # per class, in list order - the reference's classwise_wi[cls] lists concatenated
    classes = [str(c) for c in z["classes"]]
    for key in ("Abuse", "Normal", "Vandalism"):
        idx = np.concatenate([np.arange(off[v], off[v] + T[v]) for v in range(len(T)) if classes[v] == key])
        assert np.max(np.abs(res["classwise_wi"][key].cpu().numpy() - z["wi_mean"][idx])) < 1e-4
        assert res["classwise_fused"][key].shape == (idx.size, 768)
        assert torch.equal(res["classwise_fused"][key], res["fused"][torch.as_tensor(idx, device="cuda")])


@pytest.mark.gpu
def test_config2_drop_in_module_forward_all_eight_tensors_ucf_shape(gpu_model):
    """MMFMIL.forward on the [465, 256, 768] chunk batch of the bench workload: scores of all valid rows, the w_i / w_e
    row means the reference's loop derives (train/ucf_test.py:124-131) and the first row of fused / mu per video."""
    m, synth = gpu_model
    z = load_golden("config2_ucf.npz")
    m.temporal.precision = "HH"
    T = z["lengths"]
    vids = [synth.make_video(v, int(T[v])) for v in range(len(T))]
    img = torch.cat([synth.chunk_video(a) for a, _ in vids]).cuda()
    ev = torch.cat([synth.chunk_video(b) for _, b in vids]).cuda()
    assert img.shape == (465, 256, 768)
    with torch.no_grad():
        out = m(img, ev, None, None, None)
    S = np.where(T < 256, 1, T // 256 + 1)
    row0 = np.concatenate([[0], np.cumsum(S)])[:-1] * 256
    idx = torch.as_tensor(np.concatenate([np.arange(t) + r for t, r in zip(T, row0)]), device="cuda")
    sc = torch.sigmoid(out["logits"].reshape(-1)[idx]).cpu().numpy()
    assert _rel(sc, z["scores"]) < SCORE_TOL
    for k, gk in (("w_i", "wi_mean"), ("w_e", "we_mean")):
        mean = out[k].reshape(-1, 768).mean(dim=-1)[idx].cpu().numpy()
        assert np.max(np.abs(mean - z[gk])) < 1e-4, k
    r0 = torch.as_tensor(row0, device="cuda")
    for j, k in enumerate(("fused", "image_mu", "event_mu")):
        got = out[k].reshape(-1, 768)[r0].cpu().numpy()
        assert O.max_norm_err(got, z["first_rows"][:, j]) < 1e-3, k


@pytest.mark.gpu
def test_config3_xd_full_size_scores(gpu_model):
    m, synth = gpu_model
    z = load_golden("config3_xd.npz")
    m.temporal.precision = "HH"
    ev, fi, fe = _evaluator(m, synth, "xd")
    with torch.no_grad():
        ev.set_host_ragged(fi, fe)
        res = ev.step(host_inputs=True)
    sc = res["scores"].cpu().numpy()
    assert sc.size == 684801
    err = _rel(sc, z["scores"])
    assert err < SCORE_TOL, err
    assert abs(res["AUC"] - float(z["AUC"])) <= _auc_bound(z["scores"], None, err) + 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("B", [64, 128])
def test_config4_forward_and_clas2_both_batches(gpu_model, B):
    from iefvad_b200.loss import CLAS2
    m, synth = gpu_model
    m.temporal.precision = "HH"
    if B == 64:
        z = load_golden("full_default.npz")
        ref_logits, ref_loss = z["c4:logits"], float(z["c4:loss"])
    else:
        z = load_golden("config4_b128.npz")
        ref_logits, ref_loss = z["logits"], float(z["loss"])
    img, ev, lengths, labels = synth.make_c4_batch(B)
    with torch.no_grad():
        out = m(img.cuda(), ev.cuda(), None, None, lengths.cuda())
        loss = CLAS2(out["logits"], labels.cuda(), lengths.cuda(), "cuda")
    assert O.score_rel_err(out["logits"].cpu().numpy().reshape(B, 256), ref_logits) < SCORE_TOL
    assert abs(float(loss) - ref_loss) < 2e-4 * max(1.0, abs(ref_loss))
