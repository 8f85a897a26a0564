#!/bin/bash
# ncu captures for profiles/: launch list of one warm detection + match, and --set full captures of every kernel class.
# usage (on the GPU box): tools/profile_all.sh <tag>
tag=$1
# eager launches: ncu serialises kernels in CPU enqueue order, which the -s/-c windows below rely on (graph replay runs the same kernels)
export VKSIFT_GRAPH=0
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_run.py 3 2 > gpurun_out/prof_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blur_pass_fast -s 21 -c 6 -f -o gpurun_out/blur_$tag python tools/profile_run.py 2 0 >> gpurun_out/prof_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:octave_fused -s 8 -c 2 -f -o gpurun_out/fused_$tag python tools/profile_run.py 2 0 >> gpurun_out/prof_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:extrema|refine|order_prim|orientation|assemble|descriptor" -s 22 -c 22 -f -o gpurun_out/features_$tag python tools/profile_run.py 2 0 >> gpurun_out/prof_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:match_ -s 2 -c 2 -f -o gpurun_out/match_$tag python tools/profile_run.py 1 2 >> gpurun_out/prof_$tag.log 2>&1
tail -3 gpurun_out/prof_$tag.log
