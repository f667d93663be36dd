#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-700; }
run t_cfg python -m pytest tests/test_gpu_forward.py -q -x -k "c4 or c5"
for w in c1 c4 c5; do run bench_$w python bench.py --workload $w --steps 10 --warmup 3; done
run bench_xd python bench.py --workload xd --steps 5 --warmup 3 --no-cpu-baseline
