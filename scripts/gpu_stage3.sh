#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run t_ops python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py -q -x
TAILN=40 run gemm_bench python scripts/bench_gemm.py 32768
run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
