"""Multi-GPU determinism on real NCCL ranks: the gathered list-order score vector of a 2-rank (and, when
the box has them, 4-rank) sharded evaluation equals the 1-rank vector bit for bit, and so do AUC / AP.  Skipped on boxes
with fewer than 2 GPUs (the world-size-2 `gloo` test of the sharding layout, tests/test_host_logic.py, runs everywhere)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "scripts", "mgpu_determinism.py")


def _run(world, extra=()):
    if world == 1:
        cmd = [sys.executable, SCRIPT, *extra]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(29530 + world), SCRIPT, *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    return json.loads(line)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("extra", [(), ("--host-inputs",)])
def test_n_rank_scores_equal_one_rank_bit_for_bit(extra):
    one = _run(1, extra)
    worlds = [2] + ([4] if torch.cuda.device_count() >= 4 else [])
    for w in worlds:
        many = _run(w, extra)
        assert many["world"] == w and many["frames"] == one["frames"]
        assert many["sha256"] == one["sha256"], (w, one, many)
        assert many["AUC"] == one["AUC"] and many["AP"] == one["AP"] and many["ano_AUC"] == one["ano_AUC"]
