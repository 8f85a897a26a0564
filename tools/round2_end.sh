#!/bin/bash
# round-2 measurement pass on one GPU: parity tests, both bench arms, ncu launch list + full captures, sanitizer, examples, probes
tag=${1:-r2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_$tag.txt
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; tail -c 400 gpurun_out/bench_${tag}_ref.json; echo
timeout 900 python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err || tail -20 gpurun_out/bench_${tag}_n1.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_${tag}_n1.json'))
    print('feat/s', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'parity', d['parity_checked']['status'])
    print('stage_ms', {k: round(v, 4) for k, v in d.get('stage_ms', {}).items()})
    for k in ('roofline', 'roofline_stage', 'roofline_stage_pipelined'):
        if k in d: print(k, d[k].get('frac'), d[k].get('launch_us', d[k].get('stage_ms', d[k].get('marginal_ms'))))
    print('match', d['match']['ms_per_match_call'], d['match']['kernel_ms'], d['match']['roofline']['frac'], d['match']['roofline_tmem_drain']['frac'])
    print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('opencv_baseline', {}).get('value'), d['match'].get('cpu_baseline', {}).get('value'))
    print('clocks', d['clocks'])
except Exception as e:
    print('bench line unreadable:', e)
PY
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench_$tag.csv python bench.py --steps 2 --warmup 1 --quick --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
# full captures of every kernel class (eager schedule so that launch order = enqueue order)
export VKSIFT_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_detect_$tag.csv python tools/profile_run.py 3 2 > gpurun_out/prof_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blur_pass_fast -s 21 -c 6 -f -o gpurun_out/blur_$tag python tools/profile_run.py 2 0 >> gpurun_out/prof_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:extrema|refine|row_scan|rank_|orientation|assemble|descriptor" -s 24 -c 24 -f -o gpurun_out/features_$tag python tools/profile_run.py 2 0 >> gpurun_out/prof_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_ -s 2 -c 2 -f -o gpurun_out/match_$tag python tools/profile_run.py 1 2 >> gpurun_out/prof_$tag.log 2>&1
unset VKSIFT_GRAPH
# sanitizer: smoke + the lanes test
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/sanitizer_${tool}_smoke_$tag.txt 2>&1
  echo "$tool smoke: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' gpurun_out/sanitizer_${tool}_smoke_$tag.txt | tr '\n' ' ')"
done
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_detect.py -m gpu -q -k "lanes_overlap or overflow_path" > gpurun_out/sanitizer_memcheck_lanes_$tag.txt 2>&1
echo "memcheck lanes: $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitizer_memcheck_lanes_$tag.txt | tr '\n' ' ')"
# examples and probes
d=$(mktemp -d); for n in perf_runtime perf_matching; do gcc -std=c11 -O2 examples/$n.c -Iinclude -Lvulkansift_b200/lib -lvulkansift -Wl,-rpath,$PWD/vulkansift_b200/lib -lm -o $d/$n && (cd $d && timeout 300 ./$n) > gpurun_out/example_${n}_$tag.txt 2>&1; tail -3 gpurun_out/example_${n}_$tag.txt; done
timeout 300 python tools/ablation.py > gpurun_out/ablation_$tag.txt 2>&1; tail -7 gpurun_out/ablation_$tag.txt
timeout 300 python tools/fp16_probe.py > gpurun_out/fp16_probe_$tag.txt 2>&1; tail -2 gpurun_out/fp16_probe_$tag.txt
