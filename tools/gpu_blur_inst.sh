#!/bin/bash
# per-launch instruction counts and (cold) durations of the octave-0 blur launches: the quick A/B for instruction-level work
tag=$1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum -k regex:blur_pass_fast --clock-control none --csv --log-file gpurun_out/blur_inst_$tag.csv python tools/profile_run.py 2 0 > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open('gpurun_out/blur_inst_$tag.csv')) if len(r) > 10]
h = rows[0]
ki, mi, vi, gi, idi = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Grid Size'), h.index('ID')
acc = {}
for r in rows[1:]:
    acc.setdefault((int(r[idi]), r[ki][:44], r[gi]), {})[r[mi]] = float(r[vi].replace(',', ''))
tot_i = tot_t = 0
for (i, k, g), m in sorted(acc.items()):
    if '2040' in g or '(60, 34' in g:
        print(i, k, g, 'us %.2f' % (m['gpu__time_duration.sum'] / 1e3), 'Minst %.3f' % (m['smsp__inst_executed.sum'] / 1e6))
        tot_i += m['smsp__inst_executed.sum']; tot_t += m['gpu__time_duration.sum']
print('octave-0 launches of 2 detections: %.2f Minst, %.1f us' % (tot_i / 1e6, tot_t / 1e3))
PY
