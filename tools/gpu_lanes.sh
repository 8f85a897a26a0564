python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for cfg in "4 4" "2 2" "4 1"; do set -- $cfg
VKSIFT_BENCH_BUFFERS=$1 VKSIFT_LANES=$2 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_l$1_$2.json 2> gpurun_out/bench_l$1_$2.err || tail -5 gpurun_out/bench_l$1_$2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_l$1_$2.json')); print('buf $1 lanes $2: feat/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'stage_ms', {k: round(v,4) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'small', d['small_images'])"
done
