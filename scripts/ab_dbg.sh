#!/bin/bash
# timing experiments on the fused out-projection + LayerNorm kernel: scripts/ab_dbg.sh 0 1 2 3 ...  (IEFVAD_OUTPROJ_LN_DBG values)
for v in "$@"; do
  IEFVAD_OUTPROJ_LN_DBG=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-eager > gpurun_out/dbg_$v.json 2> gpurun_out/dbg_$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/dbg_$v.json"))
k = d["kernels"]
print("DBG=$v", "outproj_ln", k["outproj_ln"]["ms"], "refine", k["refine_fused"]["ms"], "ratio %.3f" % (k["outproj_ln"]["ms"] / k["refine_fused"]["ms"]))
PY
done
