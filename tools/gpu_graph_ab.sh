for g in 0 1; do
VKSIFT_GRAPH=$g python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_graph$g.json 2> gpurun_out/bench_graph$g.err || tail -5 gpurun_out/bench_graph$g.err
python -c "
import json; d=json.load(open('gpurun_out/bench_graph$g.json')); print('graph $g: feat/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'match', d['match']['ms_per_match_call'], d['match']['kernel_ms'])"
done
