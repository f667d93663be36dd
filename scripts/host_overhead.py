"""Where does a step's wall time go: host enqueue time vs device time, for the forward alone, a step without metrics and a
full asynchronous step (python scripts/host_overhead.py)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iefvad_b200 import synth
from iefvad_b200.evaluate import Evaluator
from iefvad_b200.imf_vad import MMFMIL
import bench
dev = torch.device("cuda", 0)
model = synth.build_model(MMFMIL, seed=0).to(dev).eval()
wl = bench.build_workload("ucf", 1, synth, "weak")
ev = Evaluator(model, wl["lengths"], wl["classes"], wl["gt"], device=dev)
img_c, ev_c = bench.make_features(ev, wl["video_ids"], wl["lengths"], synth, model.embed_dim)
ev.set_device_features(img_c, ev_c)
def measure(fn, n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / n * 1e3, e0.elapsed_time(e1) / n, (t2 - t0) / n * 1e3
with torch.no_grad():
    for _ in range(5): ev.step()
    for rep in range(3):
        for name, fn in (("local_scores", lambda: ev.local_scores()), ("step no metrics", lambda: ev.step(with_metrics=False, sync=False)), ("full step async", lambda: ev.step(sync=False))):
            h, d, w = measure(fn)
            print(f"{name:18s}: host enqueue {h:7.3f} ms  device {d:7.3f} ms  wall {w:7.3f} ms", flush=True)
