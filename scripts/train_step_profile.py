"""Where a training step's time goes (row N3): host enqueue vs device time of forward / backward / optimizer, launch counts,
and the per-kernel device time table (torch profiler).  python scripts/train_step_profile.py [B]"""
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import iefvad_b200  # noqa: E402
from iefvad_b200 import _lib, synth  # noqa: E402
from iefvad_b200.loss import CLAS2  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
img, ev, lengths, labels = (t.cuda() for t in synth.make_c4_batch(B))
m = synth.build_model(iefvad_b200.MMFMIL, seed=0).cuda()
m.train()
opt = torch.optim.AdamW(m.parameters(), lr=2e-5)
nu = m.temporal.nu


def fwd():      # train/ucf_train.py:60-102
    out = m(img, ev, None, None, lengths)
    mu_i, mu_e, lv_i, lv_e = out["image_mu"], out["event_mu"], out["image_logvar"], out["event_logvar"]
    loss_c = CLAS2(out["logits"], labels, lengths, img.device)
    cos = F.cosine_similarity(F.normalize(mu_i, p=2, dim=-1), F.normalize(mu_e, p=2, dim=-1), dim=-1)
    loss_reg = (1 - cos).mean() + torch.abs(torch.norm(mu_i, p=2, dim=-1) - torch.norm(mu_e, p=2, dim=-1)).mean()
    eli, ele = lv_i + math.log(nu / (nu + 1)), lv_e + math.log(nu / (nu + 1))
    return (loss_c + loss_reg - 0.5 * torch.mean(1 + eli - mu_i.pow(2) - eli.exp())
            - 0.5 * torch.mean(1 + ele - mu_e.pow(2) - ele.exp()))


for _ in range(3):
    loss = fwd(); opt.zero_grad(); loss.backward(); opt.step()
torch.cuda.synchronize()
for rep in range(2):
    l0 = _lib.lib.iefvad_launch_count()
    t0 = time.perf_counter(); loss = fwd(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    lf = _lib.lib.iefvad_launch_count() - l0
    opt.zero_grad()
    t3 = time.perf_counter(); loss.backward(); t4 = time.perf_counter(); torch.cuda.synchronize(); t5 = time.perf_counter()
    lb = _lib.lib.iefvad_launch_count() - l0 - lf
    t6 = time.perf_counter(); opt.step(); torch.cuda.synchronize(); t7 = time.perf_counter()
    print(f"forward: host {1e3 * (t1 - t0):.1f} ms, total {1e3 * (t2 - t0):.1f} ms, {lf} launches | backward: host "
          f"{1e3 * (t4 - t3):.1f}, total {1e3 * (t5 - t3):.1f}, {lb} launches | optimizer {1e3 * (t7 - t6):.1f}")
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    loss = fwd(); opt.zero_grad(); loss.backward(); opt.step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
