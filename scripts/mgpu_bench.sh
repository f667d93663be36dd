#!/bin/bash
# N-rank bench lines (weak scaling on the UCF shape, strong scaling on the XD list): scripts/mgpu_bench.sh N
N=${1:-8}
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --no-configs --no-eager "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/$name.json') if l.startswith('{')][-1]; print('$name', d['n_gpus'], d['scaling'], d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['step_breakdown'], (d.get('scores_sha256') or '')[:12], d['parity']['max_rel_err'])"; }
run r2_bench_n${N}_weak
run r2_bench_n${N}_xd_strong --workload xd --scaling strong
