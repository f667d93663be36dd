#!/bin/bash
# staged GPU bring-up: each stage in its own process under a timeout, logs into gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/$name.log; tail -n 15 gpurun_out/$name.log; }
run s1_simple python -m pytest tests/test_gpu_ops.py -q -k "fuse or layernorm or classifier" -x
run s2_linear_fp32 python -m pytest tests/test_gpu_ops.py -q -k "linear and fp32"
run s3_linear_tc python -m pytest tests/test_gpu_ops.py -q -k "linear and not fp32"
run s4_mha_fp32 python -m pytest tests/test_gpu_ops.py -q -k "mha and fp32"
run s5_mha_tc python -m pytest tests/test_gpu_ops.py -q -k "mha and bf16"
run s6_forward python -m pytest tests/test_gpu_forward.py -q
