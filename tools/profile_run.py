"""Small driver for ncu: a few C2 detections (device-resident input) and C4 matches.  Not a benchmark."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vulkansift_b200 import api
from vulkansift_b200.synth import blob_image, C2, random_descriptors

n_detect = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_match = int(sys.argv[2]) if len(sys.argv) > 2 else 2
api.lib.vksift_setLogLevel(api.VKSIFT_LOG_WARNING)
img = blob_image(**C2)
d = torch.from_numpy(img).cuda()
inst = api.Instance(max_nb_sift_per_buffer=100000)
for i in range(n_detect):
    inst.detect_device(d.data_ptr(), img.shape[1], img.shape[0], 0)
    inst.wait_idle()
print("features", inst.features_number(0))
if n_match:
    m = api.Instance(max_nb_sift_per_buffer=10000, input_image_max_size=1 << 20)
    fa = np.zeros(10000, api.FEATURE_DTYPE)
    fb = np.zeros(10000, api.FEATURE_DTYPE)
    fa["descriptor"], fb["descriptor"] = random_descriptors(10000, 1234), random_descriptors(10000, 1235)
    m.upload_features(fa, 0)
    m.upload_features(fb, 1)
    for i in range(n_match):
        m.match(0, 1)
        m.wait_idle()
    print("matches", m.matches_number())
