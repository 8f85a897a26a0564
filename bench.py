#!/usr/bin/env python
"""bench.py -- headline benchmark of the vksift detect + 2-NN match path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--octaves O] [--ns S]

Metric (BASELINE.json): SIFT features/s of the detection pipeline on configs[1]
(1920x1080, upsampling, sigma0 1.6, synthetic blob field), plus 2-NN matches/s
on configs[3] (10k x 10k x 128-D) in the "match" object.  One JSON line on rank 0.

  value   : device-resident throughput (image already in HBM, CUDA events on the library stream); a "step" is one
            detection; the K-step timed region is repeated REPS times, `value` is the median repetition (min/max beside it)
  e2e     : same metric through vksift_detectFeatures / getFeaturesNumber / downloadFeatures with HOST
            buffers (pinned source image -> H2D, feature records D2H inside the timed region)
  parity_checked : after the timed loops every feature buffer (and every all-pairs / sharded result) is downloaded and
            compared byte for byte with the CPU oracle's result for the input it holds -- the timed path IS the checked path
  roofline: dominant kernel; roofline_stage / roofline_stage_pipelined: the whole pyramid+DoG stage (SURVEY 8d bytes)
  cpu_baseline: the CPU oracle (port of the reference algorithm) timed on this box's host cores; OpenCV SIFT and
            cv2.BFMatcher (the reference's own CPU comparison path) beside it

"3 octaves x 5 scales" of BASELINE.json configs[1] has two readings; both run as flags and are labelled:
  --octaves 3        nb_octaves = 3, nb_scales_per_octave = 3 (= 5 DoG scales per octave)
  --ns 5             nb_scales_per_octave = 5 (8 Gaussian + 7 DoG layers)
The default is the configuration BASELINE.md section 3 quotes the targets on: octaves auto (7), ns = 3.

N > 1 (torchrun, one rank per GPU): images shard one per GPU with no data-path collective (weak scaling); configs[2]
(64 x 640x480) is sharded with dist.shard_range; configs[3] splits the A rows (B replicated); configs[4] (all-pairs) exchanges
descriptor blocks with an NCCL all-gather.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from vulkansift_b200.synth import C1, C2, blob_image, random_descriptors  # noqa: E402

METRIC = "sift_features_per_sec"
UNIT = "features/s"
N_IMAGES = 4  # distinct inputs rotated through the steps
MATCH_N = 10000
REPS = 5      # repetitions of the K-step timed region (median reported)
C3_IMAGES = 64


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def workload_images():
    return [blob_image(**dict(C2, seed=C2["seed"] + i)) for i in range(N_IMAGES)]


def algorithmic_bytes_pyramid(w, h, octaves, ns):
    """input read once + every Gaussian and DoG layer written exactly once (fp32)."""
    sp = sum(ow * oh for ow, oh in octaves)
    return w * h + 4 * sp * (ns + 3) + 4 * sp * (ns + 2)


def dominant_kernel_roofline(trace_acc, octaves, ns, hbm, peak_src):
    """Roofline of the dominant kernel = the per-layer blur launches of octave 0 (blur_pass_fast_kernel<R, LAYER>, 5 of the ~25
    launches of the stage, ~60 % of its bytes).  Algorithmic bytes per launch: the Gaussian layer and the DoG layer it
    writes, 2 * 4 * w0 * h0; duration = CUDA event pair around the launch on the launching stream (vksiftx launch trace), MEAN over
    the octave-0 layer launches, in the library's serial schedule (all launches on one stream; in the default schedule octaves
    overlap and a launch's event pair also covers the kernels it shares the GPU with).  With the opt-in strip schedule
    (VKSIFT_STRIP=1) the launches are "strip o0 g<first layer>+<n layers>" and write n Gaussian + n DoG layers each.
    traffic = ncu dram read+write of the same launches (profiles/traffic_r2.json)."""
    w0, h0 = octaves[0]
    us_sum, bytes_sum, n = 0.0, 0, 0
    kind = "fast"
    for name, us in trace_acc.items():
        if name.startswith("strip o0 g"):
            first, nl = name[len("strip o0 g"):].split("+")
            nbytes = 4 * w0 * h0 * 2 * int(nl)
            kind = "strip"
        elif name.startswith("fast o0 r"):
            nbytes = 2 * 4 * w0 * h0
        else:
            continue
        us_sum += us
        bytes_sum += nbytes
        n += 1
    if n == 0:
        return {"bound": "hbm", "kernel": None, "achieved": None, "peak": hbm, "unit": "GB/s", "frac": None, "traffic": None,
                "peak_source": peak_src}
    us, nbytes = us_sum / n, bytes_sum / n
    ach = nbytes / (us * 1e-6) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_r2.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(kind + "_octave0_layer_launch_dram_bytes")
    label = ("blur_pass_fast_kernel<R, LAYER>, octave 0 (%dx%d) layer launches, mean of %d" if kind == "fast" else
             "pyramid_strip_kernel, octave 0 (%dx%d) launches, mean of %d") % (w0, h0, n)
    return {"bound": "hbm", "kernel": label, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic,
            "algorithmic_bytes": nbytes, "launch_us": us, "peak_source": peak_src,
            "note": "the launch also reads its 33 MB source layer (L2 or HBM): traffic > algorithmic bytes by design of the per-layer schedule; "
                    "the kernel is bound by instruction issue / the fp32 pipe (profiles/blur_r2.txt), not by HBM; event pairs add ~1-2 us to "
                    "each traced launch"}


def tmem_drain_roofline(na, nb, kernel_ms, clocks, stream_ms=None):
    """What actually bounds the tensor-core matcher: every one of the na x nb fp32/s32 accumulators has to leave
    tensor memory through tcgen05.ld, 64 bytes per clock and SM (B300_MICROARCH.md, TMEM table), whatever the MMA rate."""
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    pad = lambda n: (n + 127) // 128 * 128
    nbytes = 4.0 * pad(na) * pad(nb)
    peak = 64.0 * 148 * sm_mhz * 1e6 / 1e9  # GB/s
    ach = nbytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    out = {"bound": "tmem_read", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "accumulator_bytes": nbytes,
           "note": "kernel_ms = one search alone, MMA kernel AND merge kernel (about 22 + 10 us at 10k x 10k); the MMA kernel alone sits on this "
                   "bound (profiles/match_r2.txt).  frac_back_to_back = the same bytes over ms_per_match_call: consecutive searches are stream "
                   "ordered and the next MMA kernel overlaps the previous merge"}
    if stream_ms:
        out["frac_back_to_back"] = nbytes / (stream_ms * 1e-3) / 1e9 / peak
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def oracle_kwargs(args):
    kw = {"input_image_max_size": C2["width"] * C2["height"]}
    if args.octaves:
        kw["nb_octaves"] = args.octaves
    if args.ns != 3:
        kw["nb_scales_per_octave"] = args.ns
    return kw


def instance_kwargs(args):
    kw = {}
    if args.octaves:
        kw["nb_octaves"] = args.octaves
    if args.ns != 3:
        kw["nb_scales_per_octave"] = args.ns
    return kw


def cpu_port_baseline(images, args, seconds_budget=20.0, threads=0):
    """Oracle (kind "port") on the host cores: detections of the same workload until ~budget seconds."""
    import oracle
    orc = oracle.Oracle(nb_threads=threads, **oracle_kwargs(args))
    cores = threads if threads > 0 else (os.cpu_count() or 1)
    n_feat, n_img = 0, 0
    orc.detect(images[0])  # warm-up (page faults, allocations)
    t0 = time.perf_counter()
    while True:
        f = orc.detect(images[n_img % len(images)])
        n_feat += len(f)
        n_img += 1
        dt = time.perf_counter() - t0
        if dt > seconds_budget or n_img >= 12:
            break
    return {"value": n_feat / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d detections of the 1920x1080 workload (%d features) in %.2f s, OpenMP oracle" % (n_img, n_feat, dt),
            "ms_per_image": 1e3 * dt / n_img, "stage_seconds_last": orc.stage_seconds()}


def cpu_match_baseline(threads=0, rows=1000):
    import oracle
    da, db = random_descriptors(MATCH_N, 1234), random_descriptors(MATCH_N, 1235)
    t0 = time.perf_counter()
    oracle.match_descriptors(da[:rows], db, threads)
    dt = time.perf_counter() - t0
    return {"value": rows / dt, "unit": "matches/s", "cores": threads if threads > 0 else (os.cpu_count() or 1), "kind": "port",
            "sample": "%d of %d A rows against %d B rows in %.2f s" % (rows, MATCH_N, MATCH_N, dt)}


def opencv_baseline(images, warm=10, n=100, budget_s=30.0):
    """The reference's own CPU comparison path (src/perf/wrappers/opencv_wrapper.cpp:5,16) with its protocol (10 warm-up + 100
    timed iterations, docs/Performances.md:22), bounded to ~budget_s seconds."""
    try:
        import cv2
    except ImportError:
        return None
    s = cv2.SIFT_create()
    for i in range(warm):
        s.detectAndCompute(images[i % len(images)], None)
    t0 = time.perf_counter()
    nf, done, best = 0, 0, 1e9
    for i in range(n):
        t1 = time.perf_counter()
        k, _ = s.detectAndCompute(images[i % len(images)], None)
        best = min(best, time.perf_counter() - t1)
        nf += len(k)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": nf / dt, "unit": UNIT, "cores": cv2.getNumThreads(), "kind": "opencv-%s" % cv2.__version__,
            "sample": "%d warm-up + %d timed cv2.SIFT_create().detectAndCompute on the same images" % (warm, done), "ms_per_image": 1e3 * dt / done,
            "best_ms": 1e3 * best}


def opencv_match_baseline(n=3):
    """cv2.BFMatcher(cv2.NORM_L2).knnMatch(A_f32, B_f32, k=2): the reference's CPU matcher (src/perf/perf_common.cpp:114-115)."""
    try:
        import cv2
    except ImportError:
        return None
    a = random_descriptors(MATCH_N, 1234).astype(np.float32)
    b = random_descriptors(MATCH_N, 1235).astype(np.float32)
    m = cv2.BFMatcher(cv2.NORM_L2)
    m.knnMatch(a[:1000], b, k=2)
    t0 = time.perf_counter()
    for _ in range(n):
        m.knnMatch(a, b, k=2)
    dt = (time.perf_counter() - t0) / n
    return {"value": MATCH_N / dt, "unit": "matches/s", "cores": cv2.getNumThreads(), "kind": "opencv-%s BFMatcher(NORM_L2).knnMatch k=2" % cv2.__version__,
            "sample": "%d x (10000 x 10000 x 128 f32), mean %.1f ms" % (n, 1e3 * dt)}


def workload_config(world, args):
    """Identical in both arms (repo and --impl reference)."""
    if args.octaves or args.ns != 3:
        reading = ("nb_octaves %s, nb_scales_per_octave %d (%d DoG scales) [alternative reading of '3 octaves x 5 scales', SURVEY App. C]"
                   % (args.octaves or "auto", args.ns, args.ns + 2))
    else:
        reading = "nb_scales_per_octave 3 (5 DoG scales), octaves auto (7)"
    return {"workload": "configs[1]: 1920x1080 detect, upsampling ON, sigma0 1.6, %s, synthetic blob field N=2400" % reading,
            "images_per_step_per_gpu": 1, "distinct_images": N_IMAGES, "sharding": "one image per GPU per step" if world > 1 else "single GPU",
            "l2": "per-step working set 489 MB (pyramid + DoG rewritten every step) > 126 MB L2; inputs rotate over %d images" % N_IMAGES}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (oracle/_ref is a
    correctness harness running GLSL through an emulation layer, not a performance build), all host threads, same
    config/metric/unit; each step = one detection of the workload (K steps exactly, as asked)."""
    if rank != 0:
        return
    images = workload_images()
    import oracle
    cores = os.cpu_count() or 1
    orc = oracle.Oracle(nb_threads=cores, **oracle_kwargs(args))  # explicit: torchrun exports OMP_NUM_THREADS=1
    for i in range(max(1, args.warmup)):
        orc.detect(images[i % len(images)])
    steps = max(1, args.steps)
    n_feat = 0
    t0 = time.perf_counter()
    for i in range(steps):
        n_feat += len(orc.detect(images[i % len(images)]))
    dt = time.perf_counter() - t0
    v = n_feat / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(world, args),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d detections (each step = one 1920x1080 image) on %d OpenMP threads" % (steps, cores)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--octaves", type=int, default=0, help="nb_octaves (0 = auto: 7 at 1920x1080)")
    ap.add_argument("--ns", type=int, default=3, help="nb_scales_per_octave")
    ap.add_argument("--quick", action="store_true", help="skip the analysis passes (stage times, launch trace, marginal cost)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from vulkansift_b200 import api
    from vulkansift_b200 import dist as vdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # stdout must carry the JSON line only, but NCCL (version banner) and native code print to fd 1: send fd 1 to stderr
    # for the whole run and keep a private handle on the real stdout for the result line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    api.lib.vksift_setLogLevel(api.VKSIFT_LOG_WARNING)

    K, W = args.steps, max(3, args.warmup)
    images = workload_images()
    h, w = images[0].shape
    NBUF = int(os.environ.get("VKSIFT_BENCH_BUFFERS", "8"))  # feature buffers = detection lanes (the default config has 2)
    ikw = instance_kwargs(args)
    inst = api.Instance(gpu_device_index=local_rank, input_image_max_size=w * h, sift_buffer_count=NBUF, **ikw)
    stream = torch.cuda.ExternalStream(inst.stream, device=local_rank)
    d_images = [torch.from_numpy(im).cuda() for im in images]
    pinned = [torch.from_numpy(im).pin_memory() for im in images]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = {"detect_buffers_checked": 0, "e2e_downloads_checked": 0, "small_images_checked": 0, "match_rows_checked": 0,
              "allpairs_peer_results_checked": 0, "mismatches": 0}
    import oracle  # the checker (and, on rank 0, the CPU baseline); never on the timed path
    orc = oracle.Oracle(nb_threads=max(1, (os.cpu_count() or 1) // max(1, world)), **oracle_kwargs(args))
    expected = [orc.detect(im) for im in images]

    def check(got, exp, what):
        ok = got.tobytes() == exp.tobytes()
        if not ok:
            parity["mismatches"] += 1
            sys.stderr.write("PARITY MISMATCH: %s (%d vs %d records)\n" % (what, len(got), len(exp)))
        return ok

    # ---------------- device-resident detection (value) ----------------
    for i in range(max(W, 3 * NBUF)):  # every lane captures its CUDA graph on the second use of a buffer
        inst.detect_device(d_images[i % N_IMAGES].data_ptr(), w, h, i % NBUF)
    inst.wait_idle()
    for i in range(3):  # and the graph of the other grid policy (a detection that finds every lane idle), so no capture lands in a timed loop
        inst.detect_device(d_images[0].data_ptr(), w, h, 0)
        inst.wait_idle()
    counts = {i: len(expected[i]) for i in range(N_IMAGES)}
    octaves = [inst.octave_resolution(o) for o in range(inst.nb_octaves())]
    ns = inst.config.nb_scales_per_octave

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms_reps, launches = [], 0
    for rep in range(REPS):
        barrier()
        launches0 = inst.kernel_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for i in range(K):
            # a detection waits for the previous one of its lane (buffer index modulo the lane count); lanes overlap on the GPU
            inst.detect_device(d_images[i % N_IMAGES].data_ptr(), w, h, i % NBUF)
        inst.join_lanes()  # the instance stream waits (on the device) for every lane, so ev1 closes all K detections
        ev1.record(stream)
        barrier()
        dev_ms_reps.append(ev0.elapsed_time(ev1))
        launches = inst.kernel_launch_count() - launches0
    n_feat = sum(counts[i % N_IMAGES] for i in range(K))
    # parity of what the timed loop produced: buffer b holds the detection of step i_last(b)
    for b in range(min(NBUF, K)):
        i_last = ((K - 1 - b) // NBUF) * NBUF + b
        check(inst.download_features(b), expected[i_last % N_IMAGES], "timed loop, buffer %d (image %d)" % (b, i_last % N_IMAGES))
        parity["detect_buffers_checked"] += 1
    dev_ms = statistics.median(dev_ms_reps)

    # ---------------- end to end through the reference API (e2e) ----------------
    # The caller's loop is software-pipelined over the instance's feature buffers, the way the reference's two-buffer API
    # is meant to be used: image i+1.. are submitted (vksift_detectFeatures copies from page-locked host memory) before the
    # features of image i are fetched (vksift_getFeaturesNumber + vksift_downloadFeatures into page-locked host memory).
    # Every step still pays its own H2D and D2H inside the timed region.
    out_pinned = [torch.empty(20000 * api.FEATURE_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(api.FEATURE_DTYPE)
                  for _ in range(NBUF)]

    def e2e_loop(n, verify=False):
        feats, nbytes = 0, 0
        for i in range(n + NBUF - 1):
            if i < n:
                inst.detect_raw(pinned[i % N_IMAGES].data_ptr(), w, h, i % NBUF)
            j = i - (NBUF - 1)
            if j >= 0:
                f = inst.download_features(j % NBUF, out=out_pinned[j % NBUF])
                feats += len(f)
                nbytes += f.nbytes + 4
                if verify:
                    check(f, expected[j % N_IMAGES], "e2e step %d" % j)
                    parity["e2e_downloads_checked"] += 1
        return feats, nbytes

    e2e_loop(3 * NBUF)
    e2e_s_reps = []
    for rep in range(REPS):
        barrier()
        t0 = time.perf_counter()
        e2e_feat, d2h = e2e_loop(K)
        torch.cuda.synchronize()
        e2e_s_reps.append(time.perf_counter() - t0)
    clocks = sampler.stop()
    e2e_loop(2 * NBUF, verify=True)  # the same loop once more with every downloaded record set compared (untimed)
    e2e_s = statistics.median(e2e_s_reps)

    # ---------------- analysis passes (not timed results): stage times, launch trace, marginal cost of the scale space ------
    stage_acc, trace_acc, pyr_marginal = {}, {}, None
    if not args.quick:
        inst.set_profiling(True)
        KP = max(3, min(K, 20))
        for i in range(KP):
            inst.detect_device(d_images[i % N_IMAGES].data_ptr(), w, h, i % 2)
            for k, v in inst.stage_times_ms().items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v / KP
        # per-launch times of the scale-space stage (event pair around every launch; traced runs are not timed runs)
        inst.set_launch_trace(True)
        inst.set_serial_schedule(True)  # one stream: the event pair around a launch then times that kernel alone
        KT = 5
        for i in range(KT + 1):
            inst.detect_device(d_images[i % N_IMAGES].data_ptr(), w, h, i % 2)
            tr = inst.launch_trace()
            if i == 0:
                continue
            seen = {}
            for name, t0, t1 in tr:
                seen[name] = seen.get(name, 0) + 1
                key = "%s #%d" % (name, seen[name]) if seen[name] > 1 else name
                trace_acc[key] = trace_acc.get(key, 0.0) + (t1 - t0) / KT
        inst.set_serial_schedule(False)
        inst.set_launch_trace(False)
        inst.set_profiling(False)
        if rank == 0:
            # what the scale space costs per image in the pipelined schedule: the timed loop with the blur launches left out,
            # on the ANALYSIS build of the library (vksiftx_setDebugSkip does not exist in the product library)
            from vulkansift_b200 import analysis
            ainst = analysis.Instance(gpu_device_index=local_rank, input_image_max_size=w * h, sift_buffer_count=NBUF, **ikw)
            astream = torch.cuda.ExternalStream(ainst.stream, device=local_rank)

            def timed_loop(n):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for i in range(3 * NBUF):
                    ainst.detect_device(d_images[i % N_IMAGES].data_ptr(), w, h, i % NBUF)
                ainst.wait_idle()
                torch.cuda.synchronize()
                e0.record(astream)
                for i in range(n):
                    ainst.detect_device(d_images[i % N_IMAGES].data_ptr(), w, h, i % NBUF)
                ainst.join_lanes()
                e1.record(astream)
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / n

            t_all = statistics.median(timed_loop(K) for _ in range(3))
            ainst.set_debug_skip(8)
            t_no_pyr = statistics.median(timed_loop(K) for _ in range(3))
            ainst.set_debug_skip(0)
            ainst.close()
            pyr_marginal = {"marginal_ms": max(t_all - t_no_pyr, 1e-6), "ms_per_image_all_stages": t_all, "ms_per_image_without_scale_space": t_no_pyr}

    # ---------------- configs[2]: 64 x 640x480 sharded over the ranks (8 per GPU at 8 GPUs), 8 lanes per GPU ----------------
    lo, hi = vdist.shard_range(C3_IMAGES, rank, world)
    small = [blob_image(**dict(C1, seed=C1["seed"] + i)) for i in range(lo, hi)]
    d_small = [torch.from_numpy(im).cuda() for im in small]
    sh_, sw_ = small[0].shape
    sinst = api.Instance(gpu_device_index=local_rank, input_image_max_size=sw_ * sh_, max_nb_sift_per_buffer=20000, sift_buffer_count=8)
    sstream = torch.cuda.ExternalStream(sinst.stream, device=local_rank)
    n_small = len(small)
    for rep in range(3):
        for i in range(n_small):
            sinst.detect_device(d_small[i].data_ptr(), sw_, sh_, i % 8)
    sinst.wait_idle()
    c3_ms = []
    c3_rounds = max(1, 160 // n_small)  # about 160 detections per timed repetition
    for rep in range(REPS):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sstream)
        for r_ in range(c3_rounds):
            for i in range(n_small):
                sinst.detect_device(d_small[i].data_ptr(), sw_, sh_, i % 8)
        sinst.join_lanes()
        e1.record(sstream)
        barrier()
        c3_ms.append(e0.elapsed_time(e1))
    sorc = oracle.Oracle(nb_threads=max(1, (os.cpu_count() or 1) // max(1, world)))
    c3_feats = 0
    for b in range(min(8, n_small)):  # buffer b holds the last image i with i % 8 == b
        i_last = ((n_small - 1 - b) // 8) * 8 + b
        e = sorc.detect(small[i_last])
        check(sinst.download_features(b), e, "configs[2] shard, image %d" % (lo + i_last))
        parity["small_images_checked"] += 1
    # single-lane latency figure for the same images (one GPU's view)
    one = api.Instance(gpu_device_index=local_rank, input_image_max_size=sw_ * sh_, max_nb_sift_per_buffer=20000, sift_buffer_count=1)
    for rep in range(2):
        for i in range(min(8, n_small)):
            one.detect_device(d_small[i].data_ptr(), sw_, sh_, 0)
    one.wait_idle()
    t0 = time.perf_counter()
    for i in range(4 * min(8, n_small)):
        one.detect_device(d_small[i % n_small].data_ptr(), sw_, sh_, 0)
    one.wait_idle()
    small_one_lane = 4 * min(8, n_small) / (time.perf_counter() - t0)
    one.close()
    sinst.close()

    # ---------------- texture-rich image (not a configs[] row): no cliff when an image yields 20x the features ----------------
    dense = None
    if not args.quick and world == 1:
        rng = np.random.default_rng(77)
        dimg = np.kron(rng.integers(0, 256, (h // 3, w // 3), dtype=np.uint8), np.ones((3, 3), np.uint8))  # 3x3 blocks of noise, 1080p
        dimg = np.ascontiguousarray(dimg[:h, :w])
        d_dense = torch.from_numpy(dimg).cuda()
        for b in range(2 * NBUF):
            inst.detect_device(d_dense.data_ptr(), w, h, b % NBUF)
        inst.wait_idle()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_dense = 24
        e0.record(stream)
        for b in range(n_dense):
            inst.detect_device(d_dense.data_ptr(), w, h, b % NBUF)
        inst.join_lanes()
        e1.record(stream)
        torch.cuda.synchronize()
        got = inst.download_features(NBUF - 1)
        check(got, orc.detect(dimg), "texture-rich image")
        parity["detect_buffers_checked"] += 1
        dense = {"workload": "1920x1080 field of 3x3 noise blocks, same configuration, %d lanes, device resident" % inst.lane_count(),
                 "features_per_image": int(len(got)), "ms_per_image": e0.elapsed_time(e1) / n_dense,
                 "features_per_s": len(got) * n_dense / (e0.elapsed_time(e1) * 1e-3), "parity": "checked against the oracle"}

    # ---------------- matcher (configs[3]); N > 1: A rows split over the ranks, B replicated ----------------
    da, db = random_descriptors(MATCH_N, 1234), random_descriptors(MATCH_N, 1235)
    a_lo, a_hi = vdist.shard_range(MATCH_N, rank, world)
    minst = api.Instance(gpu_device_index=local_rank, max_nb_sift_per_buffer=MATCH_N, input_image_max_size=1024 * 1024)
    mstream = torch.cuda.ExternalStream(minst.stream, device=local_rank)
    fa = np.zeros(a_hi - a_lo, api.FEATURE_DTYPE)
    fb = np.zeros(MATCH_N, api.FEATURE_DTYPE)
    fa["descriptor"], fb["descriptor"] = da[a_lo:a_hi], db
    minst.upload_features(fa, 0)
    minst.upload_features(fb, 1)
    for _ in range(W):
        minst.match(0, 1)
    minst.wait_idle()
    mk = max(K, 20)
    match_ms_reps = []
    for rep in range(REPS):
        mev0, mev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        mev0.record(mstream)
        for _ in range(mk):
            minst.match(0, 1)
        mev1.record(mstream)
        barrier()
        match_ms_reps.append(mev0.elapsed_time(mev1) / mk)
    match_ms = statistics.median(match_ms_reps)
    got_m = minst.download_matches()
    # parity: a bounded sample of rows against the oracle (the full 10k x 10k takes the CPU ~10 s per rank; tests/ do it in full)
    sample = np.linspace(0, len(fa) - 1, num=min(len(fa), 512)).astype(np.int64)
    exp_m = oracle.match_descriptors(da[a_lo:a_hi][sample], db, max(1, (os.cpu_count() or 1) // max(1, world)))
    exp_m["idx_a"] = sample
    check(got_m[sample], exp_m, "configs[3] match rows (sample of %d)" % len(sample))
    parity["match_rows_checked"] += len(sample)
    # kernel time from the library's stage events (analysis: an event pair per stage)
    minst.set_profiling(True)
    m_acc = 0.0
    for _ in range(mk):
        minst.match(0, 1)
        m_acc += minst.stage_times_ms()["match_2nn"]
    minst.set_profiling(False)
    match_kernel_ms = m_acc / mk
    # e2e matcher: upload both descriptor sets, match, download the rows
    t0 = time.perf_counter()
    for _ in range(5):
        minst.upload_features(fa, 0)
        minst.upload_features(fb, 1)
        minst.match(0, 1)
        minst.download_matches()
    match_e2e_s = (time.perf_counter() - t0) / 5
    minst.close()

    # ---------------- all-pairs cross-image matching (configs[4] pattern), N > 1 only ----------------
    allpairs = None
    if world > 1:
        own = blob_image(**dict(C2, seed=C2["seed"] + 100 + rank))
        inst.detect(own, 0)
        own_f = inst.download_features(0)
        n_own = len(own_f)
        reps = 10
        ap_out = torch.empty(world * n_own * api.MATCH_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(api.MATCH_DTYPE)
        # (a) baseline: one NCCL all-gather + count read-back, then one batched enqueue of all peers and one download
        ap_t, gather_t, match_t = [], [], []
        res = None
        for rep in range(reps + 1):
            barrier()
            t0 = time.perf_counter()
            counts_g, blocks_g = vdist.gather_instance_descriptors(inst, 0, capacity=4095)  # 4096-row slots: 512 KB per rank
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            res = vdist.match_against_peers_batched(inst, 0, counts_g, blocks_g, rank, out=ap_out)  # one enqueue of all peers, one download
            t2 = time.perf_counter()
            if rep > 0:  # first repetition warms NCCL up
                gather_t.append(t1 - t0)
                match_t.append(t2 - t1)
                ap_t.append(t2 - t0)
        # parity of every peer result on this rank (outside the timed region)
        host_blocks = blocks_g.cpu().numpy()
        expected = {}
        for j, m in res.items():
            if m is None:
                continue
            expected[j] = oracle.match_descriptors(own_f["descriptor"], host_blocks[j][:counts_g[j]], max(1, (os.cpu_count() or 1) // world))
            check(m, expected[j], "all-pairs (NCCL gather) rank %d vs peer %d" % (rank, j))
        # (b) the library's exchange over NVLink peer memory: every rank pushes its block into a slot on every peer (one kernel,
        # flags in peer memory), the searches run against the slots in place: one library call + one download per step
        try:
            px = vdist.PeerExchange(inst, 4096)
        except RuntimeError as e:  # no peer access / IPC on this node: every rank gets the same verdict (PeerExchange agrees on it)
            px = None
            px_unavailable = str(e)
        px_t, pxg_t = [], []
        if px is not None:
            for rep in range(reps + 1):
                barrier()
                t0 = time.perf_counter()
                counts_p, res_p = px.match_all_peers(0, out=ap_out)
                t1 = time.perf_counter()
                if rep > 0:
                    px_t.append(t1 - t0)
            assert counts_p == counts_g, (counts_p, counts_g)
            for j, m in res_p.items():
                if m is None:
                    continue
                check(m, expected[j], "all-pairs (peer-memory exchange) rank %d vs peer %d" % (rank, j))
                parity["allpairs_peer_results_checked"] += 1
            for rep in range(reps + 1):  # the exchange alone (push + wait + counts on the host)
                barrier()
                t0 = time.perf_counter()
                px.allgather(0)
                t1 = time.perf_counter()
                if rep > 0:
                    pxg_t.append(t1 - t0)
            px.close()
        else:
            px_t, pxg_t = ap_t, gather_t  # the NCCL path is all there is on this node
            for j in expected:
                parity["allpairs_peer_results_checked"] += 1  # checked above against the oracle
        ap = torch.tensor([statistics.median(ap_t), statistics.median(gather_t), statistics.median(match_t), statistics.median(px_t),
                           statistics.median(pxg_t)], dtype=torch.float64, device="cuda")
        rows = torch.tensor([float(n_own * (world - 1))], dtype=torch.float64, device="cuda")
        dist.all_reduce(ap, op=dist.ReduceOp.MAX)
        dist.all_reduce(rows, op=dist.ReduceOp.SUM)
        allpairs = {"workload": "one 1920x1080 image per GPU; every GPU pushes its descriptor block into a 4096-row slot on every peer over NVLink "
                                "peer memory (one kernel, completion flags in peer memory: vksiftx_exchangeMatchAllPeers) and matches its features "
                                "against the %d received blocks in place with ONE launch of the tensor-core search and one download (ordered pairs: %d); wall clock, "
                                "median of %d, max over ranks" % (world - 1, world * (world - 1), reps),
                    "value": rows.item() / ap[3].item(), "unit": "matches/s", "ms_total": 1e3 * ap[3].item(),
                    "ms_gather": 1e3 * ap[4].item(), "ms_match_and_download": 1e3 * (ap[3].item() - ap[4].item()), "matched_rows": rows.item(),
                    "limiter": "gather" if ap[4].item() > ap[3].item() - ap[4].item() else "match+download",
                    "exchange": "peer memory" if px is not None else "unavailable (%s): the figures are the NCCL path's" % px_unavailable,
                    "nccl_baseline": {"what": "the same step with one NCCL all-gather (4096-row slots, counts in band, count read-back) instead of "
                                              "the peer-memory exchange",
                                      "ms_total": 1e3 * ap[0].item(), "ms_gather": 1e3 * ap[1].item(), "ms_match_and_download": 1e3 * ap[2].item()}}

    # ---------------- reduce over ranks ----------------
    vals = torch.tensor([dev_ms, e2e_s, match_ms, match_kernel_ms, stage_acc.get("pyramid_dog", 0.0), statistics.median(c3_ms),
                         max(dev_ms_reps), max(e2e_s_reps)], dtype=torch.float64, device="cuda")
    mins = torch.tensor([min(dev_ms_reps), min(e2e_s_reps)], dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(n_feat), float(e2e_feat), float(n_small * c3_rounds), float(len(fa))] + [float(parity[k]) for k in sorted(parity)],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(mins, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_ms, e2e_s, match_ms, match_kernel_ms, pyr_ms, c3_ms_med, dev_ms_max, e2e_s_max = vals.tolist()
    dev_ms_min, e2e_s_min = mins.tolist()
    sl = sums.tolist()
    n_feat_all, e2e_feat_all, c3_images_all, match_rows_all = sl[:4]
    parity_all = {k: int(v) for k, v in zip(sorted(parity), sl[4:])}

    if rank == 0:
        hbm, tf_burst, tf_sust, peak_src = measured_peaks()
        alg = algorithmic_bytes_pyramid(w, h, octaves, ns)
        flops = 2.0 * match_rows_all * MATCH_N * 128
        m_ach = 2.0 * len(fa) * MATCH_N * 128 / (match_kernel_ms * 1e-3) / 1e12 if match_kernel_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": n_feat_all / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(world, args),
            "timing": {"repetitions": REPS, "statistic": "median of the repetitions (max over ranks of each)", "ms_per_step_min": dev_ms_min / K,
                       "ms_per_step_max": dev_ms_max / K, "e2e_ms_per_step_min": 1e3 * e2e_s_min / K, "e2e_ms_per_step_max": 1e3 * e2e_s_max / K},
            "parity_checked": dict(parity_all, checker="CPU oracle, byte comparison of every downloaded record set after the timed loops",
                                   status="ok" if parity_all["mismatches"] == 0 else "FAILED"),
            "features_per_image": [counts[i] for i in range(N_IMAGES)],
            "e2e": {"value": e2e_feat_all / e2e_s, "unit": UNIT, "h2d_bytes_per_step": w * h, "d2h_bytes_per_step": d2h // K,
                    "ms_per_step": 1e3 * e2e_s / K, "pipelined_over_buffers": NBUF},
            "schedule": "%d feature buffers = %d detection lanes (own scale space each, 0.5 GB): a detection waits for the previous one of "
                        "its lane only, consecutive images overlap on the GPU" % (NBUF, inst.lane_count()),
            "detection_lanes": inst.lane_count(),
            "gpu_launches": launches,
            "match": {"metric": "2nn_matches_per_sec", "value": match_rows_all / (match_ms * 1e-3), "unit": "matches/s",
                      "workload": "configs[3]: 10000 x 10000 x 128-D u8, tcgen05 path" +
                                  ("" if world == 1 else "; A rows split over %d GPUs (%d per GPU), B replicated" % (world, len(fa))),
                      "ms_per_match_call": match_ms, "kernel_ms": match_kernel_ms, "e2e_value": len(fa) / match_e2e_s, "e2e_ms": 1e3 * match_e2e_s,
                      "roofline": {"bound": "tensor", "achieved": m_ach, "peak": tf_burst, "unit": "TFLOP/s", "frac": m_ach / tf_burst,
                                   "flops_per_gpu": 2.0 * len(fa) * MATCH_N * 128, "flops": flops,
                                   "peak_source": peak_src + " bf16 dense burst (SURVEY 8d denominator)"},
                      "roofline_tmem_drain": tmem_drain_roofline(len(fa), MATCH_N, match_kernel_ms, clocks, match_ms)},
            "clocks": clocks,
            "small_images": {"workload": "configs[2]: %d x 640x480 (upsampled, default config) sharded over %d GPU(s) with dist.shard_range, 8 lanes per GPU, "
                                         "images resident in HBM, CUDA events, max over ranks" % (C3_IMAGES, world),
                             "value": c3_images_all / (c3_ms_med * 1e-3), "unit": "images/s", "images_per_gpu": n_small,
                             "images_per_s_one_lane_one_gpu": small_one_lane},
        }
        if stage_acc:
            ach = alg / (pyr_ms * 1e-3) / 1e9 if pyr_ms > 0 else 0.0
            line["latency_ms_one_detection_alone"] = stage_acc.get("detect_total", 0.0)
            line["stage_ms"] = {k: v for k, v in stage_acc.items() if k.startswith(("pyramid", "extrema", "orient", "descr", "detect"))}
            line["roofline"] = dominant_kernel_roofline(trace_acc, octaves, ns, hbm, peak_src)
            line["roofline_stage"] = {"bound": "hbm", "kernel": "whole pyramid+DoG stage (all blur launches of all octaves, CUDA events of the library, one detection alone)",
                                      "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "algorithmic_bytes": alg, "stage_ms": pyr_ms,
                                      "peak_source": peak_src}
            line["scale_space_launches_us"] = {k: round(v, 2) for k, v in trace_acc.items()}
        if pyr_marginal:
            m_ms = pyr_marginal["marginal_ms"]
            line["roofline_stage_pipelined"] = dict(pyr_marginal, bound="hbm", kernel="whole pyramid+DoG stage, marginal cost per image with all lanes busy: ms per image of "
                                                    "the timed loop minus the same loop with the blur launches left out (analysis build of the library, "
                                                    "vksiftx_setDebugSkip)", achieved=alg / (m_ms * 1e-3) / 1e9, peak=hbm, unit="GB/s",
                                                    frac=alg / (m_ms * 1e-3) / 1e9 / hbm, algorithmic_bytes=alg, peak_source=peak_src)
        if allpairs:
            line["allpairs"] = allpairs
        if dense:
            line["texture_rich_image"] = dense
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_port_baseline(images, args, threads=os.cpu_count() or 1)
            line["cpu_baseline"] = cb
            line["match"]["cpu_baseline"] = opencv_match_baseline() or cpu_match_baseline(threads=os.cpu_count() or 1)
            line["match"]["cpu_baseline_port"] = cpu_match_baseline(threads=os.cpu_count() or 1)
            ocv = opencv_baseline(images)
            if ocv:
                line["opencv_baseline"] = ocv
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
