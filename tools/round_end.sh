#!/bin/bash
# round-end measurement on the GPU box: parity tests, bench (both arms), ncu launch list of the bench command
# usage: tools/round_end.sh <tag>
tag=$1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; tail -c 600 gpurun_out/bench_${tag}_ref.json
python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err || tail -5 gpurun_out/bench_${tag}_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_${tag}_n1.json')); print('feat/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], 'clocks', d['clocks'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
tail -2 gpurun_out/ncu_bench_$tag.log
