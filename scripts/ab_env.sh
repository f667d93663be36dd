#!/bin/bash
# A/B of one environment knob on the same box: scripts/ab_env.sh IEFVAD_OUTPROJ_LN 0 1  -> step time + kernel table per value
knob=$1; shift
for v in "$@"; do
  env $knob=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs --no-eager > gpurun_out/ab_${knob}_$v.json 2> gpurun_out/ab_${knob}_$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_${knob}_$v.json"))
print("$knob=$v", "ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "err", d["parity"]["max_rel_err"])
print("   ", {k: v["ms"] for k, v in d["kernels"].items()})
PY
done
