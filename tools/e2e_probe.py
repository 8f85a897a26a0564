"""Where the end-to-end loop of bench.py loses its 5 % against the device-resident loop: the same software-pipelined loop with
the host->device copy, the download, or both taken out."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vulkansift_b200 import api
from vulkansift_b200.synth import blob_image, C2

api.load(); api.lib.vksift_setLogLevel(api.VKSIFT_LOG_WARNING)
imgs = [blob_image(**dict(C2, seed=C2["seed"] + i)) for i in range(4)]
h, w = imgs[0].shape
NB = int(os.environ.get("E2E_NB", "8"))
inst = api.Instance(input_image_max_size=w * h, sift_buffer_count=NB)
dev = [torch.from_numpy(im).cuda() for im in imgs]
pin = [torch.from_numpy(im).pin_memory() for im in imgs]
out = [torch.empty(20000 * api.FEATURE_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(api.FEATURE_DTYPE) for _ in range(NB)]

def loop(n, h2d, d2h):
    for i in range(n + NB - 1):
        if i < n:
            if h2d:
                inst.detect_raw(pin[i % 4].data_ptr(), w, h, i % NB)
            else:
                inst.detect_device(dev[i % 4].data_ptr(), w, h, i % NB)
        j = i - (NB - 1)
        if j >= 0:
            if d2h == 2:
                inst.download_features(j % NB, out=out[j % NB])
            elif d2h == 1:
                inst.features_number(j % NB)

for h2d, d2h, what in ((0, 0, "device input, no fetch"), (0, 1, "device input, count only"), (0, 2, "device input, download"),
                       (1, 0, "host input, no fetch"), (1, 1, "host input, count only"), (1, 2, "host input, download (= e2e)")):
    loop(24, h2d, d2h); inst.wait_idle()
    ts = []
    for rep in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); loop(50, h2d, d2h); inst.wait_idle(); ts.append((time.perf_counter() - t0) / 50)
    print("%-34s %.4f ms per image (min %.4f)" % (what, 1e3 * sorted(ts)[2], 1e3 * min(ts)))
