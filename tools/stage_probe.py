"""Stage times (CUDA events of the library, one detection at a time, eager schedule) of the C2 workload: analysis aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vulkansift_b200 import api  # noqa: E402
from vulkansift_b200.synth import blob_image, C2  # noqa: E402

api.load()
api.lib.vksift_setLogLevel(api.VKSIFT_LOG_WARNING)
imgs = [blob_image(**dict(C2, seed=C2["seed"] + i)) for i in range(4)]
h, w = imgs[0].shape
dev = [torch.from_numpy(im).cuda() for im in imgs]
inst = api.Instance(input_image_max_size=w * h, sift_buffer_count=1)
inst.set_profiling(True)
acc = {}
N = 40
for i in range(N + 5):
    inst.detect_device(dev[i % 4].data_ptr(), w, h, 0)
    t = inst.stage_times_ms()
    if i >= 5:
        for k, v in t.items():
            acc[k] = acc.get(k, 0.0) + v / N
print({k: round(v, 4) for k, v in acc.items() if v > 0})
