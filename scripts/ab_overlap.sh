#!/bin/bash
# N-rank A/B of the side-stream overlap of collective + ranking with the next pass: scripts/ab_overlap.sh N
N=${1:-2}
for f in "" "--no-overlap" "" "--no-overlap"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-configs --no-eager $f > gpurun_out/ov.json 2> gpurun_out/ov.err
  python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/ov.json') if l.startswith('{')][-1]; print('N=$N $f', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['step_breakdown'], d['scores_sha256'][:12])"
done
