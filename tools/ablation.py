"""What each stage costs in the pipelined (8-lane) schedule: throughput with stages left out (vksiftx_setDebugSkip of the
ANALYSIS build of the library, vulkansift_b200/analysis.py; the product library has no such switch).  Run on the GPU box."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vulkansift_b200 import analysis as api  # noqa: E402
from vulkansift_b200.synth import blob_image, C2  # noqa: E402

imgs = [blob_image(**dict(C2, seed=C2["seed"] + i)) for i in range(4)]
h, w = imgs[0].shape
dev = [torch.from_numpy(im).cuda() for im in imgs]
nbuf = 8
inst = api.Instance(input_image_max_size=w * h, sift_buffer_count=nbuf)


def run(mask, K=200):
    inst.set_debug_skip(0)
    for i in range(2 * nbuf):  # valid scale spaces and candidate lists in every lane
        inst.detect_device(dev[i % 4].data_ptr(), w, h, i % nbuf)
    inst.wait_idle()
    inst.set_debug_skip(mask)
    for i in range(3 * nbuf):
        inst.detect_device(dev[i % 4].data_ptr(), w, h, i % nbuf)
    inst.wait_idle()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        inst.detect_device(dev[i % 4].data_ptr(), w, h, i % nbuf)
    inst.wait_idle()
    return 1e3 * (time.perf_counter() - t0) / K


base = run(0)
print("all stages            %.4f ms per image" % base)
for mask, name in ((1, "descriptors"), (2, "orientation"), (3, "descriptors+orientation"), (4, "extrema+order (and what follows)"),
                   (8, "scale space"), (15, "everything (launch floor)")):
    t = run(mask)
    print("without %-28s %.4f ms  (marginal cost %.1f us)" % (name, t, 1e3 * (base - t)))
