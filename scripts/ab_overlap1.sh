#!/bin/bash
# one GPU: ranking of pass k on a side stream beside the forward of pass k + 1 (bench.py --force-overlap) vs in stream order
for f in "--force-overlap" "" "--force-overlap" ""; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs --no-eager $f > gpurun_out/ov.json 2> gpurun_out/ov.err
  python -c "
import json; d=json.load(open('gpurun_out/ov.json')); print('$f', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['scores_sha256'][:12])"
done
