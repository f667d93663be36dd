#!/bin/bash
# linear-operator tests first (each under a timeout: a protocol bug in the CTA-pair kernel traps instead of hanging)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run t_linear python -m pytest tests/test_gpu_ops.py -q -x -k "linear"
TAILN=40 run gemm_bench python scripts/bench_gemm.py 32768
