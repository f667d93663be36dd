import sys, numpy as np, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import load_golden
from oracle import iefvad_oracle as O
from iefvad_b200 import synth
from iefvad_b200.imf_vad import MMFMIL
z = load_golden("full_default.npz")
m = synth.build_model(MMFMIL, seed=0).cuda().eval()
img, ev, lengths, labels = synth.make_c4_batch()
valid = (np.arange(256)[None, :] < lengths.numpy()[:, None])
sig = lambda a: 1 / (1 + np.exp(-a.astype(np.float64)))
ref = sig(z["c4:logits"])
for plan in ("fp32", "H", "H8", "B", "A", "split", "bf16"):
    m.temporal.precision = plan
    with torch.no_grad():
        out = m(img.cuda(), ev.cuda(), None, None, None)["logits"].cpu().numpy().reshape(64, 256)
    e = np.abs(sig(out) - ref) / ref
    print(plan, "valid rows %.2e   pad rows %.2e" % (e[valid].max(), e[~valid].max()))
