# A/B of launch-shape knobs through their environment overrides (throughput of 200 pipelined 1920x1080 detections)
python -m pytest tests/test_gpu_detect.py -x -q 2>&1 | tail -2
run() { echo -n "$* : "; env "$@" python tools/lane_probe.py 4 2>&1 | tail -1 | sed 's/.*\], //'; }
run VKSIFT_EX_THREADS=256
run VKSIFT_EX_THREADS=512
run VKSIFT_DESC_CTAS=10
run VKSIFT_DESC_CTAS=12
run VKSIFT_DESC_CTAS=6
run VKSIFT_ORI_CTAS=12
run VKSIFT_ORI_CTAS=16
run VKSIFT_ORI_CTAS=4
