"""Stage times of the C2 workload with VKSIFT_PYRAMID_PRECISION_FLOAT16 (binary16 layers in HBM) next to the fp32 mode: analysis aid."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vulkansift_b200 import api  # noqa: E402
from vulkansift_b200.synth import blob_image, C2  # noqa: E402

api.load()
api.lib.vksift_setLogLevel(api.VKSIFT_LOG_WARNING)
imgs = [blob_image(**dict(C2, seed=C2["seed"] + i)) for i in range(4)]
h, w = imgs[0].shape
dev = [torch.from_numpy(im).cuda() for im in imgs]
for mode, name in ((api.VKSIFT_PYRAMID_PRECISION_FLOAT32, "fp32"), (api.VKSIFT_PYRAMID_PRECISION_FLOAT16, "fp16")):
    inst = api.Instance(input_image_max_size=w * h, sift_buffer_count=8, pyramid_precision_mode=mode)
    for i in range(24):
        inst.detect_device(dev[i % 4].data_ptr(), w, h, i % 8)
    inst.wait_idle()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 200
    for i in range(K):
        inst.detect_device(dev[i % 4].data_ptr(), w, h, i % 8)
    inst.wait_idle()
    dt = 1e3 * (time.perf_counter() - t0) / K
    one = api.Instance(input_image_max_size=w * h, sift_buffer_count=1, pyramid_precision_mode=mode)
    one.set_profiling(True)
    acc = {}
    for i in range(25):
        one.detect_device(dev[i % 4].data_ptr(), w, h, 0)
        t = one.stage_times_ms()
        if i >= 5:
            for k, v in t.items():
                acc[k] = acc.get(k, 0.0) + v / 20
    print("%s: %.4f ms per image with 8 lanes (%d features); one detection alone: %s" %
          (name, dt, inst.features_number(0), {k: round(v, 4) for k, v in acc.items() if v > 0}))
    inst.close()
    one.close()
