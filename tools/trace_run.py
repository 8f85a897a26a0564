"""Timeline of the scale-space launches of one C2 detection (VKSIFT_TRACE=1): CUDA-event start/end per launch, all streams."""
import os
import sys
os.environ["VKSIFT_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vulkansift_b200 import api
from vulkansift_b200.synth import blob_image, C2

api.lib.vksift_setLogLevel(api.VKSIFT_LOG_WARNING)
img = blob_image(**C2)
d = torch.from_numpy(img).cuda()
inst = api.Instance(max_nb_sift_per_buffer=100000)
inst.set_profiling(True)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    inst.detect_device(d.data_ptr(), img.shape[1], img.shape[0], 0)
    inst.wait_idle()
sys.stderr.flush()
t = inst.stage_times_ms()
print("stage times ms", t)
