"""Where a step's time goes: timed loop vs the per-launch event profile vs forward alone.  python scripts/step_breakdown.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from iefvad_b200 import _lib, synth  # noqa: E402
from iefvad_b200.evaluate import Evaluator  # noqa: E402
from iefvad_b200.imf_vad import MMFMIL  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda", 0)
model = synth.build_model(MMFMIL, seed=0).to(dev).eval()
wl = bench.build_workload("ucf", 0, 1, synth)
ev = Evaluator(model, wl["lengths"], wl["classes"], wl["gt"], device=dev)
img_c, ev_c = bench.make_features(ev, wl["video_ids"], wl["lengths"], synth, model.embed_dim)
ev.set_device_features(img_c, ev_c)


def timed(fn, n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    for _ in range(3):
        ev.step()
    t_step = timed(lambda: ev.step(sync=False))
    t_nomet = timed(lambda: ev.step(sync=False, with_metrics=False))
    t_fwd = timed(lambda: model.temporal.scores(ev._img, ev._ev, dev, ev._chunk_valid, ev._rowmap))
    _lib.lib.iefvad_profile_enable(1)
    ms_k, work_k, n_k = (C.c_double * 16)(), (C.c_double * 16)(), (C.c_int64 * 16)()
    tot = []
    for _ in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.temporal.scores(ev._img, ev._ev, dev, ev._chunk_valid, ev._rowmap)
        e1.record()
        torch.cuda.synchronize()
        _lib.check(_lib.lib.iefvad_profile_read(ms_k, work_k, n_k))
        tot.append((e0.elapsed_time(e1), sum(ms_k[i] for i in range(13)), sum(n_k[i] for i in range(13))))
    _lib.lib.iefvad_profile_enable(0)
    print(f"step {t_step:.3f}  step without metrics {t_nomet:.3f}  forward alone {t_fwd:.3f} ms")
    for t, s, n in tot:
        print(f"profiled forward: wall {t:.3f} ms, sum of {n} bracketed launches {s:.3f} ms")
