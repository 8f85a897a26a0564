#!/bin/bash
tag=$1
timeout 300 python -m pytest tests/test_gpu_match.py -m gpu -x -q 2>&1 | tail -15
timeout 200 python - <<'PY'
import numpy as np, torch, time, sys
sys.path.insert(0, '.')
from vulkansift_b200 import api
from vulkansift_b200.synth import random_descriptors
api.load(); api.lib.vksift_setLogLevel(api.VKSIFT_LOG_WARNING)
N = 10000
m = api.Instance(max_nb_sift_per_buffer=N, input_image_max_size=1 << 20)
fa = np.zeros(N, api.FEATURE_DTYPE); fb = np.zeros(N, api.FEATURE_DTYPE)
fa["descriptor"], fb["descriptor"] = random_descriptors(N, 1234), random_descriptors(N, 1235)
m.upload_features(fa, 0); m.upload_features(fb, 1)
st = torch.cuda.ExternalStream(m.stream)
for _ in range(10): m.match(0, 1)
m.wait_idle()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(100): m.match(0, 1)
e1.record(st); torch.cuda.synchronize()
print("ms per match call (10k x 10k):", e0.elapsed_time(e1) / 100)
m.set_profiling(True)
acc = 0
for _ in range(20):
    m.match(0, 1); acc += m.stage_times_ms()["match_2nn"]
print("kernel ms (events around tc + merge):", acc / 20)
PY
