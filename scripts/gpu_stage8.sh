#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
run t_fwd python -m pytest tests/test_gpu_forward.py tests/test_gpu_rank.py -q -x
for i in 1 2; do
run bench$i python bench.py --steps 10 --warmup 3
python - <<PY
import json
d = json.loads(open("gpurun_out/bench$i.log").readline())
print(d["ms_per_step"], d["value"], "e2e", d["e2e"], d["clocks"], d["other_plans"])
PY
done
