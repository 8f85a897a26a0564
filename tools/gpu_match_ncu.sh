#!/bin/bash
# matcher round trip: parity tests, call/kernel timing, ncu launch durations of the three matcher kernels (cold caches)
tag=$1
bash tools/gpu_match_dev.sh $tag
timeout 300 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__waves_per_multiprocessor,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/match_launches_$tag.csv python tools/profile_run.py 0 3 > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open('gpurun_out/match_launches_$tag.csv')) if len(r) > 10]
h = rows[0]
ki, mi, vi = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value')
for r in rows[1:]:
    if 'match' in r[ki] or 'norms' in r[ki]:
        print(r[ki][:40], r[mi], r[vi])
PY
