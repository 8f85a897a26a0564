"""Host-side cost of enqueueing a detection and throughput against lane count (analysis aid, run on the GPU box).

usage: python tools/lane_probe.py [buffers]   (VKSIFT_LANES / VKSIFT_GRAPH from the environment)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vulkansift_b200 import api  # noqa: E402
from vulkansift_b200.synth import blob_image, C1, C2  # noqa: E402

nbuf = int(sys.argv[1]) if len(sys.argv) > 1 else 4
if len(sys.argv) > 2 and sys.argv[2] == "small":
    C2 = C1  # 640x480 (configs[2] pattern)
api.load()
api.lib.vksift_setLogLevel(api.VKSIFT_LOG_WARNING)
imgs = [blob_image(**dict(C2, seed=C2["seed"] + i)) for i in range(4)]
h, w = imgs[0].shape
dev = [torch.from_numpy(im).cuda() for im in imgs]
inst = api.Instance(input_image_max_size=w * h, sift_buffer_count=nbuf, max_nb_sift_per_buffer=20000)
for i in range(3 * nbuf):
    inst.detect_device(dev[i % 4].data_ptr(), w, h, i % nbuf)
inst.wait_idle()
# enqueue cost: every lane idle, so the calls do not block
t = []
for i in range(inst.lane_count()):
    t0 = time.perf_counter()
    inst.detect_device(dev[i % 4].data_ptr(), w, h, i)
    t.append(1e6 * (time.perf_counter() - t0))
inst.wait_idle()
K = 200
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(K):
    inst.detect_device(dev[i % 4].data_ptr(), w, h, i % nbuf)
inst.wait_idle()
dt = time.perf_counter() - t0
print("%dx%d buffers %d lanes %d graph %s: enqueue us per call %s, %.4f ms per detection" %
      (w, h, nbuf, inst.lane_count(), os.environ.get("VKSIFT_GRAPH", "default"), [round(x, 1) for x in t], 1e3 * dt / K))
