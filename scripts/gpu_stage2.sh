#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/$name.log; tail -n 12 gpurun_out/$name.log; }
run t_all python -m pytest tests -q -m gpu -x
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench python bench.py --steps 5 --warmup 3
run bench_ref python bench.py --impl reference --steps 2 --warmup 1
run bench_A python bench.py --steps 5 --warmup 3 --plan A --no-cpu-baseline
run bench_bf16 python bench.py --steps 5 --warmup 3 --plan bf16 --no-cpu-baseline
