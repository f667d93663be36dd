"""IEFVAD_NVTX=1 wraps every stage of the forward and the evaluator's phases in NVTX ranges (tracing).
Without an attached tool the ranges are no-ops; this checks the instrumented path runs and changes no result.  Needs a B200."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = """
import hashlib, torch, iefvad_b200
from iefvad_b200 import synth
m = synth.build_model(iefvad_b200.MMFMIL, seed=0).eval().cuda()
img, ev = synth.make_video(0, 256)
with torch.no_grad():
    out = m(img[None].cuda(), ev[None].cuda(), None, None, None)
print(hashlib.sha256(out["logits"].cpu().numpy().tobytes()).hexdigest())
"""


def _run(nvtx: str) -> str:
    env = dict(os.environ, IEFVAD_NVTX=nvtx, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()[-1]


def test_nvtx_ranges_do_not_change_results():
    assert _run("1") == _run("0")
