"""Timeline of the scale-space launches of one detection (event pair per launch, all streams).
usage: python tools/trace_run.py [c1|c2] [serial]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vulkansift_b200 import api
from vulkansift_b200.synth import blob_image, C1, C2

cfg = C1 if (len(sys.argv) > 1 and sys.argv[1] == "c1") else C2
api.lib.vksift_setLogLevel(api.VKSIFT_LOG_WARNING)
img = blob_image(**cfg)
d = torch.from_numpy(img).cuda()
inst = api.Instance(max_nb_sift_per_buffer=100000, input_image_max_size=img.shape[0] * img.shape[1])
inst.set_profiling(True)
for i in range(3):
    inst.detect_device(d.data_ptr(), img.shape[1], img.shape[0], 0)
    inst.wait_idle()
print("untraced stage times ms", {k: round(v, 4) for k, v in inst.stage_times_ms().items() if v})
inst.set_launch_trace(True)
if "serial" in sys.argv:
    inst.set_serial_schedule(True)
for i in range(2):
    inst.detect_device(d.data_ptr(), img.shape[1], img.shape[0], 0)
    tr = inst.launch_trace()
for name, t0, t1 in tr:
    print("%-24s %8.1f %8.1f %7.1f" % (name, t0, t1, t1 - t0))
print("traced stage times ms", {k: round(v, 4) for k, v in inst.stage_times_ms().items() if v}, "features", inst.features_number(0))
