#!/bin/bash
# Round profile: (1) plain bench, (2) ncu launch list of the same command, (3) ncu --set full of the dominant kernel.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    $CMD > gpurun_out/prof_ncu_list.log 2>&1
echo "launch list: $(wc -l < gpurun_out/launches.csv) lines"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 60 -c 4 -f -o gpurun_out/bench_gemm_full \
    $CMD > gpurun_out/prof_ncu_full.log 2>&1
tail -2 gpurun_out/prof_ncu_full.log
