"""Where does a step's wall time go: host enqueue time vs device time, for the forward and for the metrics."""
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
model.temporal.precision = sys.argv[1] if len(sys.argv) > 1 else "B"
wl = bench.build_workload("ucf", 0, 1, synth)
ev = Evaluator(model, wl["lengths"], wl["classes"], wl["gt"], device=dev)
img_c, ev_c = bench.make_features(ev, wl["video_ids"], wl["lengths"], synth, model.embed_dim)
ev.set_device_features(img_c, ev_c)


def measure(fn, n=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e3, e0.elapsed_time(e1) / n, (t2 - t0) / n * 1e3


with torch.no_grad():
    for _ in range(3):
        ev.step()
    for rep in range(3):
        h, d, w = measure(lambda: ev.model.temporal(ev._img, ev._ev, with_scores=True))
        print(f"forward only        : host enqueue {h:8.3f} ms  device {d:8.3f} ms  wall {w:8.3f} ms")
        h, d, w = measure(lambda: ev.step(with_metrics=False))
        print(f"step without metrics: host enqueue {h:8.3f} ms  device {d:8.3f} ms  wall {w:8.3f} ms")
        h, d, w = measure(lambda: ev.step())
        print(f"full step           : host enqueue {h:8.3f} ms  device {d:8.3f} ms  wall {w:8.3f} ms")
