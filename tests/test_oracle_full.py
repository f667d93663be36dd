"""Full-size (768-d) golden checks that need no GPU: (1) our nn.Module reproduces the reference's initial weights
bit-for-bit from the seed (so the GPU box can rebuild them without the reference), (2) the numpy oracle reproduces
the reference's logits / scores / AUC on those weights."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import iefvad_oracle as O

from iefvad_b200 import synth
from iefvad_b200.imf_vad import MMFMIL


def _sd_np(model):
    return {k: v.detach().numpy() for k, v in model.state_dict().items()}


@pytest.fixture(scope="module")
def models():
    out = {}
    for tag, perturbed in (("full_default", False), ("full_perturbed", True)):
        m = synth.build_model(MMFMIL, seed=0).eval()
        if perturbed:
            synth.perturb_(m, seed=1, scale=0.1)
        out[tag] = m
    return out


@pytest.mark.parametrize("tag", ["full_default", "full_perturbed"])
def test_state_dict_digest_matches_reference(models, tag):
    z = load_golden(tag + ".npz")
    assert synth.state_digest(models[tag].state_dict()) == str(z["digest"])


def test_state_dict_keys_and_count(models):
    sd = models["full_default"].state_dict()
    assert len(sd) == 78
    assert sum(v.numel() for v in sd.values()) == 23_633_665
    assert sd["temporal.image_attn_layers.1.in_proj_weight"].shape == (2304, 768)
    assert sd["temporal.refinement_blocks.9.2.bias"].shape == (768,)
    assert sd["temporal.classifier.weight"].shape == (1, 768)


@pytest.mark.parametrize("tag", ["full_default", "full_perturbed"])
def test_oracle_full_size_c1_and_ragged(models, tag):
    z = load_golden(tag + ".npz")
    P = _sd_np(models[tag])
    img, ev = synth.make_video(0, 256)
    out = O.forward(P, img[None].numpy(), ev[None].numpy())
    assert O.score_rel_err(out["logits"].reshape(-1), z["c1:logits"]) < 2e-5
    assert O.max_norm_err(out["logits"].reshape(-1), z["c1:logits"]) < 2e-5
    rows = z["c1:rows"]
    for k in ("fused", "image_mu", "event_mu", "image_logvar", "event_logvar", "w_i", "w_e"):
        assert O.max_norm_err(out[k][0, rows], z[f"c1:{k}:rows"]) < 2e-5, k
    vids = [synth.make_video(10 + i, 40) for i in range(3)]
    img3 = torch.stack([v[0] for v in vids]).numpy()
    ev3 = torch.stack([v[1] for v in vids]).numpy()
    out = O.forward(P, img3, ev3)
    assert O.score_rel_err(out["logits"].reshape(3, 40), z["b3t40:logits"]) < 2e-5


def test_oracle_chunked_video(models):
    z = load_golden("full_default.npz")
    P = _sd_np(models["full_default"])
    img, ev = synth.make_video(20, 700)
    fi, n = O.process_split(img.numpy(), 256)
    fe, _ = O.process_split(ev.numpy(), 256)
    assert np.array_equal(fi, synth.chunk_video(img).numpy())
    out = O.forward(P, fi, fe)
    assert O.score_rel_err(out["logits"].reshape(-1)[:n], z["t700:logits"]) < 2e-5


def test_oracle_eval_loop_matches_reference_test(models):
    z = load_golden("eval_loop.npz")
    P = _sd_np(models["full_default"])
    T, classes = z["lengths"], [str(c) for c in z["classes"]]
    videos = []
    for v in range(len(T)):
        img, ev = synth.make_video(100 + v, int(T[v]))
        videos.append((img.numpy(), ev.numpy(), classes[v]))
    gt = synth.make_gt(T, classes)
    res = O.eval_loop(P, videos, gt)
    assert np.max(np.abs(res["scores"] - z["scores"]) / z["scores"]) < 2e-5
    # the ranking-based metrics move only through ties / swaps of near-equal scores
    assert abs(res["AUC"] - float(z["AUC"])) < 5e-4
    assert abs(res["AP"] - float(z["AP"])) < 5e-4
    assert abs(float(z["ret"][0]) - float(z["AUC"])) < 1e-12
    # given the reference's own scores, the restated sklearn arithmetic is exact
    rep = np.repeat(z["scores"].astype(np.float64), 16)
    assert abs(O.roc_auc_score(gt, rep) - float(z["AUC"])) < 1e-12
    assert abs(O.average_precision_score(gt, rep) - float(z["AP"])) < 1e-12


def test_oracle_c4_train_shape_forward_and_clas2(models):
    """Config 4: the reference's eval-mode forward on 64 zero-padded 256-row clips + its CLAS2 loss."""
    z = load_golden("full_default.npz")
    P = _sd_np(models["full_default"])
    img, ev, lengths, labels = synth.make_c4_batch()
    assert np.array_equal(lengths.numpy(), z["c4:lengths"]) and np.array_equal(labels.numpy(), z["c4:labels"])
    out = O.forward(P, img[:8].numpy(), ev[:8].numpy())            # batch elements are independent: 8 of 64 on the CPU
    assert O.score_rel_err(out["logits"].reshape(8, 256), z["c4:logits"][:8]) < 2e-5
    loss, _ = O.clas2(z["c4:logits"][..., None], labels.numpy(), lengths.numpy())
    assert abs(float(loss) - float(z["c4:loss"])) < 1e-6
