#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-8} gpurun_out/$name.log; }
run t_rank python -m pytest tests/test_gpu_rank.py -q -x
run host python scripts/host_overhead.py B
for i in 1 2 3; do run bench$i python bench.py --steps 10 --warmup 3 --no-cpu-baseline; done
