#!/bin/bash
# Round check: all GPU tests, smoke, default bench (+ reference arms), then the round's ncu profiles.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-5} gpurun_out/$name.log | cut -c1-600; }
run t_all python -m pytest tests -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke()"
timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "reference arm exit=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-eager"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/prof_ncu_list.log 2>&1
echo "launch list: $(wc -l < gpurun_out/launches.csv) lines"
# one --set full capture per hand-written hot kernel of the step (each from a warm pass: -s skips the first launches)
for k in refine_chain:3 outproj_ln_kernel:12 heads_fuse:3 attn_short:12 "gemm_tc:12"; do
  name=${k%%:*}; skip=${k##*:}
  ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/full_$name $CMD > gpurun_out/prof_ncu_$name.log 2>&1
  tail -1 gpurun_out/prof_ncu_$name.log
done
