"""End-to-end step time in two regimes: back-to-back steps (the copy of step k+1 overlaps step k) and isolated steps
(device idle before every step).  Usage: python scripts/e2e_isolated.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from iefvad_b200 import synth  # noqa: E402
from iefvad_b200.evaluate import Evaluator  # noqa: E402
from iefvad_b200.imf_vad import MMFMIL  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda", 0)
model = synth.build_model(MMFMIL, seed=0).to(dev).eval()
wl = bench.build_workload("ucf", 0, 1, synth)
ev = Evaluator(model, wl["lengths"], wl["classes"], wl["gt"], device=dev)
ev.set_host_ragged(*bench.make_features(ev, wl["video_ids"], wl["lengths"], synth, model.embed_dim, raw=True))
with torch.no_grad():
    for _ in range(3):
        ev.step(host_inputs=True)
    for rep in range(3):
        torch.cuda.synchronize()
        n = 15
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        t0 = time.perf_counter()
        pend = None
        evs[0].record()
        for i in range(n):
            pend = ev.step(host_inputs=True, sync=False)
            evs[i + 1].record()
        t_enq = time.perf_counter() - t0
        ev.finish(pend["pending"])
        back = (time.perf_counter() - t0) / n * 1e3
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
        iso = 0.0
        for _ in range(n):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ev.step(host_inputs=True)                      # synchronous: returns after the metrics reached the host
            iso += time.perf_counter() - t0
        print(f"plan {model.temporal.precision}: back-to-back {back:.3f} ms/step (host enqueue {t_enq / n * 1e3:.3f} ms/step)   "
              f"isolated {iso / n * 1e3:.3f} ms/step   per-step device ms: " + " ".join(f"{x:.1f}" for x in per), flush=True)
