#!/bin/bash
# ncu --set full of one GEMM micro-benchmark launch per configuration: "N K nsplit tile_n stages epi_kind"
mkdir -p gpurun_out
CFGS=${CFGS:-"768 768 3 256 0 2;768 768 1 256 0 1"}
IFS=';' read -ra LIST <<< "$CFGS"
for cfg in "${LIST[@]}"; do
  set -- $cfg
  tag="n$1_s$3_e$6"
  python scripts/one_gemm.py 32768 $cfg 3 > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed $tag"; continue; }
  ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 1 -f -o gpurun_out/gemm_$tag \
      python scripts/one_gemm.py 32768 $cfg 3 > gpurun_out/ncu_$tag.log 2>&1
  tail -1 gpurun_out/ncu_$tag.log
done
