#!/bin/bash
# Round check: all GPU tests, smoke, default bench (+ reference arm), then the round's ncu profiles.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-5} gpurun_out/$name.log | cut -c1-600; }
run t_all python -m pytest tests -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench python bench.py
run bench_ref python bench.py --impl reference --steps 2 --warmup 1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/prof_ncu_list.log 2>&1
echo "launch list: $(wc -l < gpurun_out/launches.csv) lines"
# refinement GEMMs of the 4th forward: per forward 10 encoder/head GEMMs precede the 20 refinement GEMMs
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 100 -c 4 -f -o gpurun_out/bench_gemm_refine $CMD > gpurun_out/prof_ncu_full.log 2>&1
tail -1 gpurun_out/prof_ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:attn_short -s 12 -c 1 -f -o gpurun_out/bench_attn $CMD > gpurun_out/prof_ncu_attn.log 2>&1
tail -1 gpurun_out/prof_ncu_attn.log
