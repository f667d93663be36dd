"""Drop-in MMFMIL.forward on the UCF chunk batch [465, 256, 768]: refinement chain fused vs per-step GEMMs, per-kernel profile."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from iefvad_b200 import _lib, synth
from iefvad_b200.imf_vad import MMFMIL

dev = torch.device("cuda:0")
model = synth.build_model(MMFMIL, seed=0).to(dev).eval()
T = synth.config_lengths("ucf")
vids = [synth.make_video(v, int(T[v])) for v in range(len(T))]
img = torch.cat([synth.chunk_video(a) for a, _ in vids]).to(dev)
ev = torch.cat([synth.chunk_video(b) for _, b in vids]).to(dev)
with torch.no_grad():
    for mode in (False, True, None):
        model.temporal.refine_fused = mode
        for _ in range(3):
            model(img, ev, None, None, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            model(img, ev, None, None, None)
        e1.record()
        torch.cuda.synchronize()
        _lib.lib.iefvad_profile_enable(1)
        model(img, ev, None, None, None)
        torch.cuda.synchronize()
        ms, work, n = bench.read_kernel_profile(_lib)
        _lib.lib.iefvad_profile_enable(0)
        print("refine_fused =", mode, "ms/forward", round(e0.elapsed_time(e1) / 5, 3),
              {k: round(ms[i], 3) for i, k in enumerate(bench.KERNEL_CLASSES) if n[i]})
