#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
run t_fwd python -m pytest tests/test_gpu_forward.py -q -x
for plan in B H; do
run bench_$plan python bench.py --steps 10 --warmup 3 --no-cpu-baseline --plan $plan
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$plan.log").readline())
print("$plan", d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["frame_auc"])
for k, v in d["kernels"].items(): print(" ", k, v)
PY
done
python - <<'PY'
# score error of each plan vs the fp32 CUDA plan on a few chunks of the bench workload
import sys, torch, numpy as np
sys.path.insert(0, ".")
from iefvad_b200 import synth
from iefvad_b200.imf_vad import MMFMIL
m = synth.build_model(MMFMIL, seed=0).cuda().eval()
img, ev = synth.make_video(3, 2000)
ci, ce = synth.chunk_video(img).cuda(), synth.chunk_video(ev).cuda()
outs = {}
with torch.no_grad():
    for plan in ("fp32", "B", "A", "H", "bf16"):
        m.temporal.precision = plan
        outs[plan] = torch.sigmoid(m(ci, ce, None, None, None)["logits"].double()).cpu().numpy().reshape(-1)
for plan in ("B", "A", "H", "bf16"):
    print("plan", plan, "max rel score err vs fp32 plan: %.2e" % np.max(np.abs(outs[plan] - outs["fp32"]) / outs["fp32"]))
PY
