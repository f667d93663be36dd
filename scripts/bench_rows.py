"""Device timings of the SURVEY 8(d) rows that the headline bench does not break out: AUC / AP ranking (benchmark size
and the 2^24 stress point), MIL top-k loss, stand-alone fusion, and the dead-class rows D1-D3 at T = 16384.
Usage (GPU box): python scripts/bench_rows.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from iefvad_b200 import layers, ops  # noqa: E402

HBM = 6451.0        # GB/s, MEASURED_PEAKS.json fallback
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as fh:
        pk = json.load(fh)
        HBM = float(pk.get("hbm_gbs", pk.get("hbm_copy_gbs", HBM)))
except Exception:
    pass


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
g = torch.Generator(device="cuda").manual_seed(0)
with torch.no_grad():
    for n in (77788, 150000, 1 << 24):
        s = torch.rand(n, device="cuda", generator=g)
        pos = (torch.rand(n, device="cuda", generator=g) < 0.05).to(torch.int32) * 16
        ms = timed(lambda: ops.auc_ap(s, pos), n=5 if n > 1e6 else 20)
        out[f"auc_ap n={n}"] = {"ms": round(ms, 4), "algorithmic_GBps": round(24.0 * n / ms / 1e6, 1),
                                "frac_of_hbm": round(24.0 * n / ms / 1e6 / HBM, 4),
                                "lsd_sort_min_traffic_GBps": round(64.0 * n / ms / 1e6, 1)}
    for B in (64, 128):
        logits = torch.randn(B, 256, 1, device="cuda", generator=g)
        labels = torch.zeros(B, 14, device="cuda")
        labels[: B // 2, 0] = 1
        labels[B // 2:, 1] = 1
        lengths = torch.randint(16, 257, (B,), device="cuda", generator=g)
        out[f"clas2 B={B}"] = {"us": round(1e3 * timed(lambda: ops.clas2(logits, labels, lengths), n=50), 2)}
    M = 77788
    t = [torch.randn(M, 768, device="cuda", generator=g) for _ in range(4)]
    ms = timed(lambda: ops.fuse(*t))
    out["fuse M=77788 (all three outputs)"] = {"ms": round(ms, 4), "GBps": round(M * 768 * 28 / ms / 1e6, 1),
                                              "frac_of_hbm": round(M * 768 * 28 / ms / 1e6 / HBM, 4)}
    T = 16384
    x = torch.randn(1, T, 768, device="cuda", generator=g)
    da = layers.DistanceAdj().cuda()
    ms = timed(lambda: da(1, T), n=5)
    out["DistanceAdj T=16384"] = {"ms": round(ms, 3), "GBps": round(4.0 * T * T / ms / 1e6, 1),
                                  "frac_of_hbm": round(4.0 * T * T / ms / 1e6 / HBM, 4)}
    sa = layers.SimilarityAdj(768, 768).cuda()
    ms = timed(lambda: sa(x, None), n=5)
    out["SimilarityAdj T=16384"] = {"ms": round(ms, 3), "write_GBps": round(4.0 * T * T / ms / 1e6, 1),
                                    "TFLOPs": round((2.0 * T * 768 * 768 + 2.0 * T * T * 768) / ms / 1e9, 1)}
    gc = layers.GraphConvolution(768, 768, residual=True).cuda()
    adj = da(1, T)
    ms = timed(lambda: gc(x, adj), n=5)
    out["GraphConvolution dense T=16384"] = {"ms": round(ms, 3),
                                             "TFLOPs": round((2.0 * T * 768 * 768 + 2.0 * T * T * 768) / ms / 1e9, 1)}
    ms = timed(lambda: gc(x, None), n=10)
    out["GraphConvolution with DistanceAdj as bidirectional scan T=16384"] = {"ms": round(ms, 3)}
    sup = torch.randn(1, T, 768, device="cuda", generator=g)
    ms = timed(lambda: layers.distance_scan(sup), n=20)
    out["distance_scan T=16384"] = {"ms": round(ms, 4), "GBps": round(3.0 * T * 768 * 4 / ms / 1e6, 1)}
print(json.dumps(out, indent=1))
