#!/bin/bash
# quick GPU round trip: full gpu test suite + short bench + (optional) ablation table
tag=$1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -20 gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$tag.json'))
    print('feat/s', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), d['timing']['ms_per_step_min'], d['timing']['ms_per_step_max'], 'e2e', round(d['e2e']['value']), 'parity', d['parity_checked']['status'], d['parity_checked']['mismatches'])
    print('stage_ms', {k: round(v, 4) for k, v in d.get('stage_ms', {}).items()})
    for k in ('roofline', 'roofline_stage', 'roofline_stage_pipelined'):
        if k in d: print(k, d[k].get('frac'), d[k].get('launch_us', d[k].get('stage_ms', d[k].get('marginal_ms'))))
    print('match', d['match']['ms_per_match_call'], d['match']['kernel_ms'], d['match']['roofline']['frac'])
    print('small', d['small_images']['value'], 'launches', d['gpu_launches'])
except Exception as e:
    print('bench line unreadable:', e)
PY
if [ "$2" == "ablation" ]; then timeout 300 python tools/ablation.py 2>&1 | tail -9; fi
