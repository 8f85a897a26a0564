#!/bin/bash
# two-GPU round trip: gpu tests on one GPU, then the bench line at N=2 (all-pairs and sharded sections included)
tag=$1
N=${2:-2}
if [ "$3" != "notests" ]; then timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 --quick > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err || tail -30 gpurun_out/bench_${tag}_n$N.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_${tag}_n$N.json'))
    print('N', d['n_gpus'], 'feat/s', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'parity', d['parity_checked'])
    print('match', d['match']['value'], d['match']['ms_per_match_call'], d['match']['workload'])
    print('small', d['small_images'])
    print('allpairs', d.get('allpairs'))
except Exception as e:
    print('bench line unreadable:', e)
PY
