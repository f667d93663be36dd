"""Deterministic synthetic workloads of the five BASELINE.json configs.

Everything is generated on the CPU generator so that the authoring container, the GPU box and the committed
golden fixtures see bit-identical weights and inputs.  No reference code and no oracle is imported here."""
from __future__ import annotations

import hashlib
from types import SimpleNamespace
from typing import Dict, List, Tuple

import numpy as np
import torch

UCF_CLASSES = ["Abuse", "Arrest", "Arson", "Assault", "Burglary", "Explosion", "Fighting", "RoadAccidents",
               "Robbery", "Shooting", "Shoplifting", "Stealing", "Vandalism", "Normal"]   # train/ucf_test.py:33-39
XD_CLASSES = ["normal", "fighting", "shooting", "riot", "abuse", "car accident", "explosion"]  # train/ucf_test.py:40-45


def default_args(**over) -> SimpleNamespace:
    """The six args.* fields MMFMIL reads (model/imf_vad.py:30-38) at their main.py:71-77 defaults."""
    a = dict(visual_layers=2, visual_head=8, num_refinement_steps=10, lambda_ref=0.5, noise_model="StudentT", nu=8)
    a.update(over)
    return SimpleNamespace(**a)


def build_model(cls, seed: int = 0, embed_dim: int = 768, args=None, device="cpu"):
    """cls(14, D, 256, D, 8, 2, 8, 10, 10, device, args) under torch.manual_seed(seed) - works for the reference
    class and for ours (identical RNG consumption order)."""
    args = args or default_args()
    torch.manual_seed(seed)
    return cls(14, embed_dim, 256, embed_dim, 8, 2, 8, 10, 10, device, args)


def perturb_(model: torch.nn.Module, seed: int = 1, scale: float = 0.1) -> None:
    """Default init leaves MHA biases 0 and LN affine (1, 0), which would hide bias / affine bugs: add
    scale * randn to every bias and every LayerNorm weight (deterministic, in state_dict order)."""
    g = torch.Generator("cpu").manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("bias") or "norms." in name or "whiten_" in name or "ln_" in name:
                p.add_(scale * torch.randn(p.shape, generator=g).to(p.device))


def state_digest(sd: Dict[str, torch.Tensor]) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def make_video(v: int, T: int, D: int = 768, dtype=torch.float16) -> Tuple[torch.Tensor, torch.Tensor]:
    """Image / event embeddings of video v: unit-direction randn rows scaled to L2 norm 10 (the CLIP ViT-L/14
    embedding scale), seeds 1000+v / 2000+v, stored in `dtype` like the .npy files."""
    out = []
    for base in (1000, 2000):
        g = torch.Generator("cpu").manual_seed(base + v)
        x = torch.randn(T, D, generator=g)
        x = 10.0 * x / x.norm(dim=1, keepdim=True)
        out.append(x.to(dtype))
    return out[0], out[1]


def chunk_video(feat: torch.Tensor, length: int = 256) -> torch.Tensor:
    """data/tools.py:100-114 (`process_split`) followed by the caller's unsqueeze (train/ucf_test.py:79-81):
    always returns [S, length, D]; T < length -> one zero-padded chunk, else int(T/length)+1 chunks with the last
    zero-padded (an all-zero extra chunk when T % length == 0)."""
    T, D = feat.shape
    S = 1 if T < length else int(T / length) + 1
    out = feat.new_zeros((S, length, D))
    out.view(S * length, D)[:T] = feat
    return out


def config_lengths(name: str) -> np.ndarray:
    """Per-video embedding-row counts T_v of configs C2 ('ucf') and C3 ('xd')."""
    if name == "ucf":
        rng = np.random.default_rng(2)
        T = np.clip(np.round(np.exp(rng.normal(np.log(180.0), 0.9, 290))), 1, 4096).astype(np.int64)
        T[:8] = [1, 15, 16, 255, 256, 257, 512, 4096]
        return T
    if name == "xd":
        rng = np.random.default_rng(3)
        return np.clip(np.round(np.exp(rng.normal(np.log(512.0), 1.0, 800))), 32, 8192).astype(np.int64)
    raise ValueError(name)


def config_classes(name: str, n: int) -> List[str]:
    """Round-robin class names so every class key has >= 1 video (train/ucf_test.py:166 needs that);
    'ucf' puts the normal class on every other video like the real test list (150 of 290 normal)."""
    if name == "ucf":
        abn = UCF_CLASSES[:-1]
        return [("Normal" if i % 2 else abn[(i // 2) % len(abn)]) for i in range(n)]
    ab = XD_CLASSES[1:]
    return [("normal" if i % 2 else ab[(i // 2) % len(ab)]) for i in range(n)]


def make_gt(lengths, classes) -> np.ndarray:
    """Frame-level ground truth, 16 raw frames per embedding row (list/ucf_generate_gt.py:24): abnormal videos
    get one positive interval [4 T_v, 9 T_v) of their 16 T_v frames."""
    parts = []
    for T, c in zip(lengths, classes):
        g = np.zeros(16 * int(T), dtype=np.float64)
        if c not in ("Normal", "normal"):
            g[4 * int(T):9 * int(T)] = 1.0
        parts.append(g)
    return np.concatenate(parts) if parts else np.zeros(0)


def make_c4_batch(B: int = 64, T: int = 256, seed: int = 4, num_class: int = 14):
    """Config 4 (train/ucf_train.py:44-73 shape): B zero-padded clips of T rows, lengths ~ U{16..T}, first half
    normal.  -> (img [B,T,D] fp16, ev [B,T,D] fp16, lengths int64 [B], labels fp32 [B, num_class])."""
    rng = np.random.default_rng(seed)
    lengths = rng.integers(16, T + 1, B)
    clips = [make_video(100 + i, T) for i in range(B)]
    img = torch.stack([c[0] for c in clips])
    ev = torch.stack([c[1] for c in clips])
    for i in range(B):                                        # process_feat zero-pads short clips (data/tools.py:89-97)
        img[i, int(lengths[i]):] = 0
        ev[i, int(lengths[i]):] = 0
    labels = torch.zeros(B, num_class)
    labels[: B // 2, 0] = 1.0
    labels[B // 2:, 1 + (torch.arange(B - B // 2) % (num_class - 1))] = 1.0
    return img, ev, torch.from_numpy(lengths.astype(np.int64)), labels


def make_event_frames(name: str) -> np.ndarray:
    """Seeded uint8 frame stacks [B, 16, 224, 224, 3] for the event-synthesis row (extracting/ucf_gen_event.py:85-91):
    `rand` = independent noise, `smooth` = a fixed picture plus sigma-14 noise (frame differences straddle the
    thresholds 10 / 25 the extraction scripts use), `static` = constant frames (no event anywhere: 0 / 0 = NaN)."""
    rng = np.random.default_rng({"rand": 21, "smooth": 22, "static": 23, "ties": 24}[name])
    if name == "rand":
        return rng.integers(0, 256, (2, 16, 224, 224, 3), dtype=np.uint8)
    if name == "smooth":
        base = rng.integers(0, 256, (2, 1, 224, 224, 3))
        return np.clip(base + rng.normal(0, 14, (2, 16, 224, 224, 3)), 0, 255).astype(np.uint8)
    if name == "static":
        return np.full((1, 16, 224, 224, 3), 77, dtype=np.uint8)
    if name == "ties":
        # every pixel pair whose exact gray difference lies within 4e-5 of the thresholds 25 / 10 (integer channel
        # differences a, b, c with 0.2989 a + 0.587 b + 0.114 c ~ threshold): the comparison `diff > threshold` is then
        # decided by the last bit of the fp32 evaluation order - these cases pin that order.  [1, 2, 1, N, 3]
        rows = []
        for thr in (25.0, 10.0):
            for a in range(-255, 256):
                for b in range(-255, 256):
                    c = (thr - 0.2989 * a - 0.587 * b) / 0.114
                    for cc in (int(np.floor(c)), int(np.ceil(c))):
                        if -255 <= cc <= 255 and abs(0.2989 * a + 0.587 * b + 0.114 * cc - thr) < 4e-5:
                            for base in range(0, 256, 17):
                                p1 = (base + a, base + b, base + cc)
                                if all(0 <= x <= 255 for x in p1):
                                    rows.append(((base, base, base), p1))
        rows = rows[: len(rows) // 4 * 4]
        f0 = np.array([r[0] for r in rows], dtype=np.uint8)
        f1 = np.array([r[1] for r in rows], dtype=np.uint8)
        return np.stack([f0, f1], 0)[None, :, None, :, :].copy()
    raise KeyError(name)


def make_locmap_case(n_videos: int = 40, seed: int = 31, num_class: int = 14):
    """Synthetic inputs of the localisation-mAP row (train/metrics.py:44-136): per-video [T, 14] class predictions with a
    few raised plateaus (so that thresholding yields proposals), and ground-truth segments that partly overlap them.
    -> (predictions: list of float32 arrays, gtsegments, gtlabels)."""
    classlist = ['Normal', 'Abuse', 'Arrest', 'Arson', 'Assault', 'Burglary', 'Explosion', 'Fighting', 'RoadAccidents',
                 'Robbery', 'Shooting', 'Shoplifting', 'Stealing', 'Vandalism']
    rng = np.random.default_rng(seed)
    preds, segs, labels = [], [], []
    for v in range(n_videos):
        T = int(rng.integers(20, 400))
        p = (0.05 * rng.standard_normal((T, num_class)) + 0.1).astype(np.float32)
        p[:, rng.integers(0, num_class, 2)] -= 0.5                        # a couple of classes with a non-positive score
        vs, vl = [], []
        for _ in range(int(rng.integers(1, 4))):
            c = int(rng.integers(0, num_class))
            a = int(rng.integers(0, T - 8))
            b = min(T, a + int(rng.integers(4, max(5, T // 3))))
            p[a:b, c] += np.float32(rng.uniform(0.3, 0.9))
            shift = int(rng.integers(-6, 7))
            vs.append([max(0, a + shift), max(1, b + shift)])
            vl.append(classlist[c])
        preds.append(p)
        segs.append(vs)
        labels.append(vl)
    for c, name in enumerate(classlist):                                  # every class gets at least one proposal + one gt
        v = c % n_videos
        T = preds[v].shape[0]
        preds[v][2:9, c] += np.float32(1.5)
        preds[v][:, c] += np.float32(0.6)
        segs[v].append([1, 10])
        labels[v].append(name)
    return preds, segs, labels
