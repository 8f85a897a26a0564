#!/bin/bash
# one GPU round trip: parity tests, short bench, ncu launch list, optional full capture of the blur kernels
# usage: tools/gpu_cycle.sh <tag> [full]
tag=$1
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); print('feat/s', round(d['value']), 'stage_ms', {k: round(v,4) for k,v in d['stage_ms'].items()}, 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), 'match', d['match']['kernel_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_run.py 3 0 > gpurun_out/prof_$tag.log 2>&1
if [ "$2" == "full" ]; then
  ncu --set full --clock-control none --import-source on -k regex:blur_pass_fast -s 11 -c 11 -f -o gpurun_out/blur_$tag python tools/profile_run.py 2 0 >> gpurun_out/prof_$tag.log 2>&1
fi
tail -2 gpurun_out/prof_$tag.log
