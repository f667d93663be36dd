"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
(/root/reference, read-only) and scikit-learn in the authoring container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed here as small .npz files and are
what `oracle/iefvad_oracle.py` (and, through it and directly, the CUDA path) is pinned against.  Weights and
inputs of the full-size cases are regenerated from seeds by `ief-vad_b200/synth.py`; the fixtures store a sha256 of
the state_dict so a seed/RNG drift is detected instead of silently comparing different models."""
import os
import sys
import types

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("IEFVAD_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# stubs for packages the reference's eval loop imports but that are not installed / not wanted
for name in ("matplotlib", "matplotlib.pyplot", "wandb"):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.log = lambda *a, **k: None
        m.init = lambda *a, **k: None
        sys.modules[name] = m
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

from model.imf_vad import MMFMIL as RefMMFMIL  # noqa: E402  (reference)
from model import layers as ref_layers  # noqa: E402
from model import module as ref_module  # noqa: E402
from train.loss import CLAS2 as ref_CLAS2  # noqa: E402
from train import ucf_test as ref_ucf_test  # noqa: E402
from sklearn.metrics import average_precision_score, roc_auc_score  # noqa: E402

import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "ief-vad_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

torch.set_num_threads(8)
OUT_KEYS = ["fused", "logits", "image_mu", "event_mu", "image_logvar", "event_logvar", "w_i", "w_e"]


def sd_np(model):
    return {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def small_models():
    """Tiny models with every tensor stored: weights, inputs, all 8 outputs."""
    cases = [
        ("small_studentt", dict(noise_model="StudentT", nu=8, num_refinement_steps=3, visual_head=4), 128, 2, 24),
        ("small_gaussian", dict(noise_model="Gaussian", nu=8, num_refinement_steps=2, visual_head=2), 128, 3, 17),
        ("small_r0", dict(noise_model="StudentT", nu=5, num_refinement_steps=0, visual_head=4, lambda_ref=0.25),
         128, 1, 40),
    ]
    for name, over, D, B, T in cases:
        args = synth.default_args(**over)
        model = synth.build_model(RefMMFMIL, seed=7, embed_dim=D, args=args).eval()
        synth.perturb_(model, seed=11, scale=0.2)
        g = torch.Generator("cpu").manual_seed(5)
        img = torch.randn(B, T, D, generator=g)
        ev = torch.randn(B, T, D, generator=g)
        with torch.no_grad():
            out = model(img, ev, None, None, None)
        arrays = {"param:" + k: v for k, v in sd_np(model).items()}
        arrays.update({"out:" + k: out[k].numpy() for k in OUT_KEYS})
        arrays.update(img=img.numpy(), ev=ev.numpy(), heads=np.int64(args.visual_head), nu=np.float64(args.nu),
                      lambda_ref=np.float64(args.lambda_ref), noise_model=np.array(args.noise_model))
        save(name + ".npz", **arrays)


def full_models():
    """Full-size (768-d, 2 layers, 8 heads, 10 refinement steps) model at the default seed: logits + row samples of
    the wide tensors + digests.  Weights/inputs come from synth (seed 0 / video seeds)."""
    rows = np.array([0, 1, 37, 100, 255])
    for tag, perturbed in (("full_default", False), ("full_perturbed", True)):
        model = synth.build_model(RefMMFMIL, seed=0).eval()
        if perturbed:
            synth.perturb_(model, seed=1, scale=0.1)
        digest = synth.state_digest(model.state_dict())
        arrays = {"digest": np.array(digest)}
        # C1: B=1, T=256 (fp16 storage like the .npy embeddings)
        img, ev = synth.make_video(0, 256)
        with torch.no_grad():
            out = model(img[None], ev[None], None, None, None)
        arrays["c1:logits"] = out["logits"].numpy().reshape(-1)
        for k in OUT_KEYS:
            if k != "logits":
                arrays[f"c1:{k}:rows"] = out[k].numpy()[0, rows]
        arrays["c1:rows"] = rows
        # ragged direct call: B=3, T=40 (T not a multiple of the 128-row tiles, 3 batch elements in one tile)
        vids = [synth.make_video(10 + i, 40) for i in range(3)]
        img3 = torch.stack([v[0] for v in vids])
        ev3 = torch.stack([v[1] for v in vids])
        with torch.no_grad():
            out = model(img3, ev3, None, None, None)
        arrays["b3t40:logits"] = out["logits"].numpy().reshape(3, 40)
        for k in OUT_KEYS:
            if k != "logits":
                arrays[f"b3t40:{k}:row7"] = out[k].numpy()[:, 7]
        # a chunked video: T=700 -> 3 chunks of 256 (data/tools.py:100-114), fp32 logits of the 700 valid rows
        img7, ev7 = synth.make_video(20, 700)
        ci, ce = synth.chunk_video(img7), synth.chunk_video(ev7)
        with torch.no_grad():
            out = model(ci, ce, None, None, None)
        arrays["t700:logits"] = out["logits"].numpy().reshape(-1)[:700]
        # direct long call T=1000 (multi key-block streaming softmax, ragged tail)
        img1k, ev1k = synth.make_video(21, 1000)
        with torch.no_grad():
            out = model(img1k[None], ev1k[None], None, None, None)
        arrays["t1000:logits"] = out["logits"].numpy().reshape(-1)
        # C4 (ucf_train.py shape): B=64 clips of T=256 (train/ucf_train.py:44-48 feeds pooled / padded 256-row clips),
        # eval-mode forward + the reference's CLAS2 (train/loss.py:18-30) with lengths ~ U{16..256}, half normal labels
        rng = np.random.default_rng(4)
        B4 = 64
        lengths = rng.integers(16, 257, B4)
        clips = [synth.make_video(100 + i, 256) for i in range(B4)]
        img4 = torch.stack([c[0] for c in clips])
        ev4 = torch.stack([c[1] for c in clips])
        for i in range(B4):                                    # process_feat zero-pads short clips (data/tools.py:89-97)
            img4[i, int(lengths[i]):] = 0
            ev4[i, int(lengths[i]):] = 0
        labels = torch.zeros(B4, 14)
        labels[: B4 // 2, 0] = 1.0                               # first half normal (column 0), rest abnormal
        labels[B4 // 2:, 1 + (torch.arange(B4 // 2) % 13)] = 1.0
        with torch.no_grad():
            out = model(img4, ev4, None, None, torch.from_numpy(lengths))
            loss = ref_CLAS2(out["logits"], labels, torch.from_numpy(lengths), "cpu")
        arrays["c4:lengths"] = lengths.astype(np.int64)
        arrays["c4:labels"] = labels.numpy()
        arrays["c4:logits"] = out["logits"].numpy().reshape(B4, 256)
        arrays["c4:loss"] = np.float64(float(loss))
        # C5: one video of T=16384 fed directly (no chunking): attention over all 16384 keys
        if os.environ.get("IEFVAD_GOLDEN_SKIP_C5") != "1":
            img5, ev5 = synth.make_video(30, 16384)
            with torch.no_grad():
                out = model(img5[None], ev5[None], None, None, None)
            arrays["c5:logits"] = out["logits"].numpy().reshape(-1)
            arrays["c5:fused:rows"] = out["fused"].numpy()[0, rows * 64]
        save(tag + ".npz", **arrays)


def clas2_cases():
    g = torch.Generator("cpu").manual_seed(3)
    B, T = 16, 256
    logits = 2.0 * torch.randn(B, T, 1, generator=g)
    # force ties and extreme values
    logits[0, :40] = 0.5
    logits[1, 3] = 30.0
    logits[2] = -30.0
    lengths = torch.tensor([256, 255, 17, 16, 15, 1, 100, 200, 32, 31, 33, 64, 128, 250, 2, 240])
    labels = torch.zeros(B, 14)
    labels[::2, 0] = 1.0
    labels[1::2, 3] = 1.0
    loss = ref_CLAS2(logits, labels, lengths, "cpu")
    save("clas2.npz", logits=logits.numpy(), lengths=lengths.numpy(), labels=labels.numpy(), loss=loss.numpy())


def eval_loop_case():
    """The reference's own train/ucf_test.py:test() on a synthetic loader: 30 UCF-style videos, full-size
    default-seed model, CPU."""
    n = 30
    rng = np.random.default_rng(9)
    T = np.clip(np.round(np.exp(rng.normal(np.log(120.0), 0.9, n))), 1, 900).astype(np.int64)
    T[:6] = [1, 16, 255, 256, 257, 512]
    classes = synth.config_classes("ucf", n)
    gt = synth.make_gt(T, classes)
    model = synth.build_model(RefMMFMIL, seed=0).eval()

    class Loader:
        def __iter__(self):
            for v in range(n):
                img, ev = synth.make_video(100 + v, int(T[v]))
                # what data/dataset.py:34-52 + DataLoader(batch_size=1) deliver
                from data.tools import process_split
                fi, ln = process_split(img.numpy(), 256)
                fe, _ = process_split(ev.numpy(), 256)
                yield (torch.from_numpy(fi)[None], torch.from_numpy(fe)[None], [classes[v]], torch.tensor([ln]))

    args = types.SimpleNamespace(exp_name="golden", dataset="ucfcrime")
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        captured = {}
        real_auc = ref_ucf_test.roc_auc_score

        def spy(gt_, pred):
            captured.setdefault("first", np.asarray(pred)[::16].copy())
            return real_auc(gt_, pred)

        ref_ucf_test.roc_auc_score = spy
        ret = ref_ucf_test.test(args, model, Loader(), 256, None, gt, "cpu")
        ref_ucf_test.roc_auc_score = real_auc
    finally:
        os.chdir(cwd)
    scores = captured["first"]
    rep = np.repeat(scores, 16)
    auc, ap = roc_auc_score(gt, rep), average_precision_score(gt, rep)
    print("reference test() returned", ret, "recomputed", auc, ap)
    save("eval_loop.npz", lengths=T, classes=np.array(classes), scores=scores.astype(np.float32), AUC=np.float64(auc),
         AP=np.float64(ap), ret=np.array([float(x) for x in ret]) if ret is not None else np.zeros(0))


class _Recorder(torch.nn.Module):
    """Wraps the unmodified reference model and records, per call, what the reference's eval loop derives from the
    forward besides the scores (train/ucf_test.py:124-144): w_i.mean(-1), w_e.mean(-1) and sample rows of fused /
    image_mu / event_mu (the full [len, 768] tensors would be 240 MB per tensor on config 2)."""

    def __init__(self, model):
        super().__init__()
        self.model = model
        self.wi, self.we, self.rows = [], [], []
        self.len_cur = 0             # set by the loader: the loop's own `lengths` over-counts when T % 256 == 0

    def forward(self, img, ev, padding_mask, text, lengths):
        out = self.model(img, ev, padding_mask, text, lengths)
        n = self.len_cur
        flat = lambda t: t.reshape(t.shape[0] * t.shape[1], t.shape[2])  # noqa: E731
        self.wi.append(flat(out["w_i"]).mean(dim=-1).numpy()[:n].copy())
        self.we.append(flat(out["w_e"]).mean(dim=-1).numpy()[:n].copy())
        self.rows.append(np.stack([flat(out[k]).numpy()[0].copy() for k in ("fused", "image_mu", "event_mu")]))
        return out


def config2_case():
    """BASELINE configs[1] AT FULL SIZE (the bench workload): the reference's own train/ucf_test.py:test() over all 290
    synthetic UCF-shaped videos (77 788 frames, T up to 4096 = 17 chunks), default-seed full-size model, CPU fp32."""
    T = synth.config_lengths("ucf")
    n = len(T)
    classes = synth.config_classes("ucf", n)
    gt = synth.make_gt(T, classes)
    model = _Recorder(synth.build_model(RefMMFMIL, seed=0).eval())
    digest = synth.state_digest(model.model.state_dict())
    from data.tools import process_split

    class Loader:
        def __iter__(self):
            for v in range(n):
                img, ev = synth.make_video(v, int(T[v]))
                fi, ln = process_split(img.numpy(), 256)
                fe, _ = process_split(ev.numpy(), 256)
                model.len_cur = int(ln)
                yield (torch.from_numpy(fi)[None], torch.from_numpy(fe)[None], [classes[v]], torch.tensor([ln]))

    args = types.SimpleNamespace(exp_name="golden", dataset="ucfcrime")
    cwd = os.getcwd()
    os.chdir("/tmp")
    captured = {}
    real_auc = ref_ucf_test.roc_auc_score

    def spy(gt_, pred):
        captured.setdefault("first", np.asarray(pred)[::16].copy())
        return real_auc(gt_, pred)

    try:
        ref_ucf_test.roc_auc_score = spy
        ret = ref_ucf_test.test(args, model, Loader(), 256, None, gt, "cpu")
    finally:
        ref_ucf_test.roc_auc_score = real_auc
        os.chdir(cwd)
    scores = captured["first"]
    assert scores.size == int(T.sum())
    rep = np.repeat(scores, 16)
    auc, ap = roc_auc_score(gt, rep), average_precision_score(gt, rep)
    # class-wise AUC / AP (:164-178) and Ano-AUC (:336-353) recomputed with sklearn on the captured scores
    off = np.concatenate([[0], np.cumsum(T)])
    keys = list(dict.fromkeys(classes))
    cw = np.full((len(keys), 2), np.nan)
    for c, key in enumerate(keys):
        idx = np.concatenate([np.arange(off[v], off[v + 1]) for v in range(n) if classes[v] == key])
        g = gt.reshape(-1, 16)[idx].reshape(-1)
        if g.sum() == 0:
            continue
        r = np.repeat(scores[idx], 16)
        cw[c] = roc_auc_score(g, r), average_precision_score(g, r)
    idx = np.concatenate([np.arange(off[v], off[v + 1]) for v in range(n) if classes[v] != "Normal"])
    ano = roc_auc_score(gt.reshape(-1, 16)[idx].reshape(-1), np.repeat(scores[idx], 16))
    print("config 2: reference test() returned", ret, "recomputed", auc, ap, "ano", ano)
    save("config2_ucf.npz", digest=np.array(digest), lengths=T, classes=np.array(classes),
         scores=scores.astype(np.float32), AUC=np.float64(auc), AP=np.float64(ap), ano_AUC=np.float64(ano),
         class_keys=np.array(keys), classwise=cw, wi_mean=np.concatenate(model.wi).astype(np.float32),
         we_mean=np.concatenate(model.we).astype(np.float32), first_rows=np.stack(model.rows).astype(np.float32),
         ret=np.array([float(x) for x in ret]) if ret is not None else np.zeros(0))


def config3_case():
    """BASELINE configs[2] AT FULL SIZE: 800 synthetic XD-Violence-shaped videos (684 801 frames) through the
    per-video loop of train/xd_test.py:70-119 (the same loop as ucf_test: process_split chunks -> model ->
    sigmoid(logits[:len])); scores + overall AUC / AP with sklearn."""
    T = synth.config_lengths("xd")
    n = len(T)
    classes = synth.config_classes("xd", n)
    gt = synth.make_gt(T, classes)
    model = synth.build_model(RefMMFMIL, seed=0).eval()
    from data.tools import process_split
    out_scores = []
    for v in range(n):
        img, ev = synth.make_video(v, int(T[v]))
        fi, ln = process_split(img.numpy(), 256)
        fe, _ = process_split(ev.numpy(), 256)
        fi, fe = torch.from_numpy(fi), torch.from_numpy(fe)
        if ln < 256:
            fi, fe = fi[None], fe[None]
        out = model(fi, fe, None, None, None)
        lg = out["logits"].reshape(-1, 1)
        out_scores.append(torch.sigmoid(lg[0:ln].squeeze(-1)).numpy())
    scores = np.concatenate(out_scores)
    rep = np.repeat(scores.astype(np.float64), 16)
    auc, ap = roc_auc_score(gt, rep), average_precision_score(gt, rep)
    print("config 3:", scores.size, "frames, AUC", auc, "AP", ap)
    save("config3_xd.npz", digest=np.array(synth.state_digest(model.state_dict())), lengths=T,
         classes=np.array(classes), scores=scores.astype(np.float32), AUC=np.float64(auc), AP=np.float64(ap))


def config4_b128_case():
    """Config 4 at the batch train/ucf_train.py:44-48 really builds (two 64-clip loaders concatenated -> B = 128):
    eval-mode forward + the reference's CLAS2, inputs from synth.make_c4_batch(128)."""
    model = synth.build_model(RefMMFMIL, seed=0).eval()
    img, ev, lengths, labels = synth.make_c4_batch(128)
    out = model(img, ev, None, None, lengths)
    loss = ref_CLAS2(out["logits"], labels, lengths, "cpu")
    save("config4_b128.npz", digest=np.array(synth.state_digest(model.state_dict())), lengths=lengths.numpy(),
         labels=labels.numpy(), logits=out["logits"].numpy().reshape(128, 256), loss=np.float64(float(loss)))


def _reference_train_loss(model, img, ev, lengths, labels, nu):
    """The loss of train/ucf_train.py:60-102 (classification + cos / norm regulariser + StudentT KL), verbatim."""
    import math
    import torch.nn.functional as F
    outputs = model(img, ev, None, None, lengths)
    logits = outputs['logits']
    image_mu, event_mu = outputs['image_mu'], outputs['event_mu']
    image_logvar, event_logvar = outputs['image_logvar'], outputs['event_logvar']
    loss_classification = ref_CLAS2(logits, labels, lengths, "cpu")
    image_mu_norm = F.normalize(image_mu, p=2, dim=-1)
    event_mu_norm = F.normalize(event_mu, p=2, dim=-1)
    cos_sim = F.cosine_similarity(image_mu_norm, event_mu_norm, dim=-1)
    loss_cos = 1 - cos_sim
    norm_image = torch.norm(image_mu, p=2, dim=-1)
    norm_event = torch.norm(event_mu, p=2, dim=-1)
    loss_norm = torch.abs(norm_image - norm_event)
    loss_reg = loss_cos.mean() + loss_norm.mean()
    effective_logvar_image = image_logvar + math.log(nu / (nu + 1))
    effective_logvar_event = event_logvar + math.log(nu / (nu + 1))
    kl_loss_image = -0.5 * torch.mean(1 + effective_logvar_image - image_mu.pow(2) - effective_logvar_image.exp())
    kl_loss_event = -0.5 * torch.mean(1 + effective_logvar_event - event_mu.pow(2) - effective_logvar_event.exp())
    loss_kl = kl_loss_image + kl_loss_event
    return loss_classification + 1 * loss_reg + 1 * loss_kl, loss_classification


def train_step_cases():
    """Gradients of the reference's training loss (train/ucf_train.py:60-105) by the reference's own autograd, eval mode
    (attention dropout off - its mask comes from PyTorch's generator, which no other implementation can reproduce).
    small: every gradient tensor of a 128-d model; full: the 768-d default-seed model on 8 clips of config 4 - gradient
    norms plus 64 sampled entries per parameter (the full set would be 94 MB)."""
    with torch.enable_grad():
        args = synth.default_args(noise_model="StudentT", nu=8, num_refinement_steps=3, visual_head=4)
        model = synth.build_model(RefMMFMIL, seed=7, embed_dim=128, args=args).eval()
        synth.perturb_(model, seed=11, scale=0.2)
        g = torch.Generator("cpu").manual_seed(5)
        B, T = 3, 40
        img, ev = torch.randn(B, T, 128, generator=g), torch.randn(B, T, 128, generator=g)
        lengths = torch.tensor([40, 17, 33])
        labels = torch.zeros(B, 14)
        labels[0, 0] = 1.0
        labels[1, 3] = 1.0
        labels[2, 5] = 1.0
        loss, lcls = _reference_train_loss(model, img, ev, lengths, labels, 8)
        loss.backward()
        arrays = {"param:" + k: v for k, v in sd_np(model).items()}
        arrays.update({"grad:" + k: p.grad.numpy().copy() for k, p in model.named_parameters()})
        arrays.update(img=img.numpy(), ev=ev.numpy(), lengths=lengths.numpy(), labels=labels.numpy(),
                      loss=np.float64(float(loss)), loss_cls=np.float64(float(lcls)), heads=np.int64(4), nu=np.float64(8))
        save("train_small.npz", **arrays)

        model = synth.build_model(RefMMFMIL, seed=0).eval()
        img, ev, lengths, labels = synth.make_c4_batch(8)
        loss, lcls = _reference_train_loss(model, img, ev, lengths, labels, 8)
        loss.backward()
        rng = np.random.default_rng(17)
        arrays = {"digest": np.array(synth.state_digest(model.state_dict())), "loss": np.float64(float(loss)),
                  "loss_cls": np.float64(float(lcls)), "lengths": lengths.numpy(), "labels": labels.numpy()}
        for k, p in model.named_parameters():
            gnp = p.grad.numpy().reshape(-1)
            idx = np.sort(rng.choice(gnp.size, size=min(64, gnp.size), replace=False))
            arrays["idx:" + k] = idx.astype(np.int64)
            arrays["val:" + k] = gnp[idx].copy()
            arrays["norm:" + k] = np.float64(np.linalg.norm(gnp.astype(np.float64)))
        save("train_full.npz", **arrays)


def locmap_case():
    """train/metrics.py run as-is: getDetectionMAP on synthetic class predictions (synth.make_locmap_case)."""
    from train import metrics as ref_metrics
    import warnings
    preds, segs, labels = synth.make_locmap_case()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dmap, ious = ref_metrics.getDetectionMAP([p.copy() for p in preds], segs, labels, excludeNormal=False)
        short = ref_metrics.getLocMAP([p.copy() for p in preds[:9]], 0.3, segs[:9], labels[:9], False)     # a class without proposals -> 0
    print("locmap:", dmap, ious, "first 9 videos:", short)
    save("locmap.npz", dmap=np.array(dmap, dtype=np.float64), ious=np.array(ious), short=np.float64(short))


def sklearn_cases():
    rng = np.random.default_rng(12)
    arrays = {}
    for i, (n, tie, posr) in enumerate([(1000, 0.0, 0.3), (5000, 0.4, 0.05), (64, 0.9, 0.5), (3, 0.0, 0.5),
                                        (2000, 0.2, 1.0), (2000, 0.2, 0.0)]):
        s = rng.random(n).astype(np.float32)
        if tie > 0:
            s = np.round(s * (1.0 / tie)) * tie
            s = s.astype(np.float32)
        pos = (rng.random(n) < posr)
        cnt = np.where(pos, rng.integers(1, 17, n), 0).astype(np.int64)     # positives among the 16 frames
        gt = np.zeros((n, 16))
        for j in range(n):
            gt[j, :cnt[j]] = 1
        gt = gt.reshape(-1)
        rep = np.repeat(s.astype(np.float64), 16)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                auc = roc_auc_score(gt, rep)
            except ValueError:
                auc = float("nan")
            ap = average_precision_score(gt, rep)
        arrays[f"{i}:scores"] = s
        arrays[f"{i}:pos"] = cnt
        arrays[f"{i}:auc"] = np.float64(auc)
        arrays[f"{i}:ap"] = np.float64(ap)
    arrays["n"] = np.int64(6)
    save("sklearn_auc.npz", **arrays)


def layer_cases():
    torch.manual_seed(21)
    arrays = {}
    B, T, Din, Dout = 2, 50, 128, 128
    x = torch.randn(B, T, Din)
    sim = ref_layers.SimilarityAdj(Din, Dout)
    with torch.no_grad():
        arrays["sim:w0"] = sim.weight0.numpy().copy()
        arrays["sim:x"] = x.numpy()
        arrays["sim:out_none"] = sim(x, None).numpy()
        arrays["sim:out_len"] = sim(x, [50, 31]).numpy()
        arrays["sim:seq_len"] = np.array([50, 31])
        gc = ref_layers.GraphConvolution(Din, Dout, bias=True, residual=True)
        adj = torch.softmax(torch.randn(B, T, T), dim=-1)
        arrays["gc:w"] = gc.weight.numpy().copy()
        arrays["gc:b"] = gc.bias.numpy().copy()
        arrays["gc:adj"] = adj.numpy()
        arrays["gc:out"] = gc(x, adj).numpy()
        gc2 = ref_layers.GraphConvolution(Din, 256, bias=False, residual=True)
        arrays["gc2:w"] = gc2.weight.numpy().copy()
        arrays["gc2:conv_w"] = gc2.residual.weight.numpy().copy()
        arrays["gc2:conv_b"] = gc2.residual.bias.numpy().copy()
        arrays["gc2:out"] = gc2(x, adj).numpy()
        # Transformer (seq-first), 2 layers, with key padding mask and additive attn mask
        W, Hh, Lyr, Lseq, N = 128, 4, 2, 40, 3
        mask = torch.zeros(Lseq, Lseq)
        mask[torch.triu(torch.ones(Lseq, Lseq), diagonal=9) > 0] = -1e4
        tr = ref_module.Transformer(W, Lyr, Hh, attn_mask=mask).eval()
        synth.perturb_(tr, seed=4, scale=0.2)
        xt = torch.randn(Lseq, N, W)
        pad = torch.zeros(N, Lseq, dtype=torch.bool)
        pad[1, 30:] = True
        pad[2, 5:] = True
        out_masked, _ = tr((xt, pad))
        out_plain, _ = tr((xt, None))
        for k, v in tr.state_dict().items():
            arrays["tr:param:" + k] = v.numpy().copy()
        arrays["tr:x"] = xt.numpy()
        arrays["tr:pad"] = pad.numpy()
        arrays["tr:mask"] = mask.numpy()
        arrays["tr:out_masked"] = out_masked.numpy()
        arrays["tr:out_nopad"] = out_plain.numpy()
        arrays["tr:heads"] = np.int64(Hh)
    save("layers.npz", **arrays)


def tools_cases():
    """data/tools.py:65-114 run as-is: process_split / process_feat (uniform_extract, pad) on seeded ragged inputs."""
    from data import tools as ref_tools
    arrays = {}
    lens = [1, 15, 100, 255, 256, 257, 300, 511, 512, 513, 1000, 4097]
    arrays["lens"] = np.array(lens)
    for dt_name, dt in (("f32", np.float32), ("f16", np.float16)):
        rng = np.random.default_rng(12)
        for t in lens:
            feat = rng.standard_normal((t, 16)).astype(dt)
            sp, n = ref_tools.process_split(feat, 256)
            arrays[f"{dt_name}:{t}:split"] = np.asarray(sp)
            arrays[f"{dt_name}:{t}:split_len"] = np.int64(n)
            pf, m = ref_tools.process_feat(feat, 256)
            arrays[f"{dt_name}:{t}:feat"] = np.asarray(pf)
            arrays[f"{dt_name}:{t}:feat_len"] = np.int64(m)
    save("tools.npz", **arrays)


def event_cases():
    """extracting/ucf_gen_event.py: `generate_event_image` (:21-37) run as-is plus the clamp / normalise / stack lines of
    its caller (:91-95, transcribed below with their line numbers).  The module itself imports cv2 / clip / matplotlib
    at the top, none of which is installed, so only the function's own source is executed (ast-extracted, unmodified)."""
    import ast
    path = os.path.join(REF, "extracting", "ucf_gen_event.py")
    src = open(path).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "generate_event_image")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    generate_event_image = ns["generate_event_image"]
    import hashlib
    arrays = {}
    for name in ("rand", "smooth", "static", "ties"):
        frames = synth.make_event_frames(name)
        arrays[f"{name}:sha256"] = np.array(hashlib.sha256(frames.tobytes()).hexdigest())
        for thr, clamp in ((25, 10), (10, 10), (25, 3)):
            event = generate_event_image(frames, thr)                 # :91  [B, H, W] counts
            arrays[f"{name}:{thr}:sum"] = event.numpy().astype(np.uint8)          # counts <= 15
            assert np.array_equal(arrays[f"{name}:{thr}:sum"].astype(np.float32), event.numpy())
            event = torch.clamp(event, 0, clamp)                      # :92
            if event.numel() != 0:
                event = event / event.max()                           # :93-94
            event = torch.stack([event, event, event], 1)             # :95
            arrays[f"{name}:{thr}:{clamp}:event"] = event.numpy()[:, 0].copy()    # the three channels are one tensor
    save("event.npz", **arrays)


if __name__ == "__main__":
    which = sys.argv[1:] or ["small", "full", "clas2", "eval", "sklearn", "layers", "tools", "event"]
    with torch.no_grad():
        if "small" in which:
            small_models()
        if "full" in which:
            full_models()
        if "clas2" in which:
            clas2_cases()
        if "eval" in which:
            eval_loop_case()
        if "sklearn" in which:
            sklearn_cases()
        if "layers" in which:
            layer_cases()
        if "tools" in which:
            tools_cases()
        if "event" in which:
            event_cases()
        if "config2" in which:            # full-size cases: not in the default list (about 20 s / 3 min of CPU)
            config2_case()
        if "config3" in which:
            config3_case()
        if "config4" in which:
            config4_b128_case()
        if "train" in which:
            train_step_cases()
        if "locmap" in which:
            locmap_case()
