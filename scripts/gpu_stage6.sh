#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
run t_fwd python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py tests/test_gpu_layers.py -q -x -k "mha or forward or full or small or batch or transformer or dtypes or load_state"
run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").readline())
print(d["ms_per_step"], d["e2e"]["ms_per_step"])
for k, v in d["kernels"].items(): print(" ", k, v)
print(d["roofline"])
PY
