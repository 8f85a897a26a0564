#!/bin/bash
# development round trip for the strip kernel: the pyramid parity tests first (with a hang guard), then the rest
tag=$1
timeout 300 python -m pytest tests/test_gpu_detect.py -m gpu -x -q -k "pyramid_and_dog or c2_full or ragged or alternative or fp16" 2>&1 | tail -15
echo "--- full suite"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -20 gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$tag.json'))
    print('feat/s', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'parity', d['parity_checked']['status'], d['parity_checked']['mismatches'])
    print('stage_ms', {k: round(v, 4) for k, v in d.get('stage_ms', {}).items()})
    for k in ('roofline', 'roofline_stage', 'roofline_stage_pipelined'):
        if k in d: print(k, d[k].get('frac'), d[k].get('launch_us', d[k].get('stage_ms', d[k].get('marginal_ms'))))
    print('match', d['match']['ms_per_match_call'], d['match']['kernel_ms'], d['match']['roofline']['frac'])
    print('small', d['small_images']['value'], 'launches', d['gpu_launches'])
    print('trace', d.get('scale_space_launches_us'))
except Exception as e:
    print('bench line unreadable:', e)
PY
