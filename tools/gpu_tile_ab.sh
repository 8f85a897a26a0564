#!/bin/bash
# A/B of blur tile heights: detect parity tests + three lane-probe runs (8 lanes, 1080p)
timeout 400 python -m pytest tests/test_gpu_detect.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2 3; do timeout 200 python tools/lane_probe.py 8 2>&1 | tail -1 | sed 's/.*\], //'; done
