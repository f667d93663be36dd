"""Evaluation step with resident inputs: eager launches vs one CUDA-graph replay per step.
Usage: python scripts/graph_step.py [workload]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from iefvad_b200 import synth  # noqa: E402
from iefvad_b200.evaluate import Evaluator  # noqa: E402
from iefvad_b200.imf_vad import MMFMIL  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda", 0)
model = synth.build_model(MMFMIL, seed=0).to(dev).eval()
wl = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else "ucf", 0, 1, synth)
ev = Evaluator(model, wl["lengths"], wl["classes"], wl["gt"], device=dev)
img_c, ev_c = bench.make_features(ev, wl["video_ids"], wl["lengths"], synth, model.embed_dim)
ev.set_device_features(img_c, ev_c)


def timed(fn, n=30):
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
    eager = timed(lambda: ev.step(sync=False))
    eager_nometrics = timed(lambda: ev.step(sync=False, with_metrics=False))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            ev.step(sync=False)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        res = ev.step(sync=False)
    graph = timed(g.replay)
    out = ev.finish(res["pending"])
    ref = ev.step()
    print(f"eager {eager:.3f} ms/step   eager without metrics {eager_nometrics:.3f}   graph replay {graph:.3f} ms/step   "
          f"AUC graph {out['AUC']:.12f} eager {ref['AUC']:.12f}")
