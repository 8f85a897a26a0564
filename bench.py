#!/usr/bin/env python
"""bench.py -- headline benchmark of the vksift detect + 2-NN match path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): SIFT features/s of the detection pipeline on configs[1]
(1920x1080, upsampling, sigma0 1.6, synthetic blob field), plus 2-NN matches/s
on configs[3] (10k x 10k x 128-D) in the "match" object.  One JSON line on rank 0.

  value   : device-resident throughput (image already in HBM, CUDA events on the library stream)
  e2e     : same metric through vksift_detectFeatures / getFeaturesNumber / downloadFeatures with HOST
            buffers (pinned source image -> H2D, feature records D2H inside the timed region)
  roofline: pyramid+DoG stage, algorithmic bytes / CUDA-event stage time / measured HBM peak
  cpu_baseline: the CPU oracle (port of the reference algorithm) timed on this box's host cores

N > 1 (torchrun, one rank per GPU): images shard one per GPU with no data-path collective
(weak scaling); the all-pairs match step exchanges descriptor blocks with an NCCL all-gather.
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

from vulkansift_b200.synth import C2, blob_image, random_descriptors  # noqa: E402

METRIC = "sift_features_per_sec"
UNIT = "features/s"
N_IMAGES = 4  # distinct inputs rotated through the steps
MATCH_N = 10000


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


def dominant_kernel_roofline(trace_acc, octaves, hbm, peak_src):
    """Roofline of the dominant kernel = the per-layer blur launches of octave 0 (blur_pass_fast_kernel<R, LAYER>, 5 of the
    ~25 launches, ~60 % of the stage's bytes).  Algorithmic bytes per launch: the Gaussian layer and the DoG layer
    it writes, 2 * 4 * w0 * h0; duration = CUDA event pair around the launch on the launching stream (vksiftx launch trace),
    averaged over the octave-0 layer launches, taken in the library's serial schedule (all launches on one stream; in the
    default schedule octaves overlap and a launch's event pair also covers the kernels it shares the GPU with).  traffic = ncu dram read+write of the same kernel, profiles/traffic_r1.json."""
    w0, h0 = octaves[0]
    per_launch = 2 * 4 * w0 * h0
    durs = [v for k, v in trace_acc.items() if k.startswith("fast o0 r") and not k.endswith("#1")]
    if not durs:
        return {"bound": "hbm", "kernel": "blur_pass_fast_kernel (octave 0 layer launches)", "achieved": None, "peak": hbm, "unit": "GB/s",
                "frac": None, "traffic": None, "peak_source": peak_src}
    us = sum(durs) / len(durs)
    ach = per_launch / (us * 1e-6) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_r1.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("blur_pass_fast_kernel_octave0_layer_bytes_per_launch")
    return {"bound": "hbm", "kernel": "blur_pass_fast_kernel<R, LAYER>, octave 0 (3840x2160) layer launches, mean of %d" % len(durs),
            "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic, "algorithmic_bytes": per_launch,
            "launch_us": us, "peak_source": peak_src,
            "note": "the launch also reads its 33 MB source layer (L2 or HBM); event pairs add ~1-2 us to each traced launch"}


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


def cpu_port_baseline(images, seconds_budget=20.0, threads=0):
    """Oracle (kind "port") on the host cores: detections of the same workload until ~budget seconds."""
    import oracle
    orc = oracle.Oracle(nb_threads=threads)
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


def opencv_baseline(images, n=3):
    """The reference's own CPU comparison path (src/perf/wrappers/opencv_wrapper.cpp:5,16), reported beside the port."""
    try:
        import cv2
    except ImportError:
        return None
    s = cv2.SIFT_create()
    s.detectAndCompute(images[0], None)
    t0 = time.perf_counter()
    nf = 0
    for i in range(n):
        k, _ = s.detectAndCompute(images[i % len(images)], None)
        nf += len(k)
    dt = time.perf_counter() - t0
    return {"value": nf / dt, "unit": UNIT, "cores": cv2.getNumThreads(), "kind": "opencv-%s" % cv2.__version__,
            "sample": "%d x cv2.SIFT_create().detectAndCompute on the same images" % n, "ms_per_image": 1e3 * dt / n}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (oracle/_ref is a
    correctness harness, not a performance build), all host threads, same config/metric/unit."""
    if rank != 0:
        return
    images = workload_images()
    import oracle
    cores = os.cpu_count() or 1
    orc = oracle.Oracle(nb_threads=cores)  # explicit: torchrun exports OMP_NUM_THREADS=1
    for _ in range(max(1, min(args.warmup, 2))):
        orc.detect(images[0])
    steps = max(1, min(args.steps, 10))
    n_feat = 0
    t0 = time.perf_counter()
    for i in range(steps):
        n_feat += len(orc.detect(images[i % len(images)]))
    dt = time.perf_counter() - t0
    v = n_feat / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(world),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d detections (each step = one 1920x1080 image) on %d OpenMP threads" % (steps, cores)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(world):
    return {"workload": "configs[1]: 1920x1080 detect, upsampling ON, sigma0 1.6, nb_scales_per_octave 3 (5 DoG scales), "
                        "octaves auto (7), synthetic blob field N=2400",
            "images_per_step_per_gpu": 1, "distinct_images": N_IMAGES, "sharding": "one image per GPU per step" if world > 1 else "single GPU",
            "l2": "per-step working set 489 MB (pyramid + DoG rewritten every step) > 126 MB L2; inputs rotate over %d images" % N_IMAGES}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    inst = api.Instance(gpu_device_index=local_rank, input_image_max_size=w * h, sift_buffer_count=NBUF)
    stream = torch.cuda.ExternalStream(inst.stream, device=local_rank)
    d_images = [torch.from_numpy(im).cuda() for im in images]
    pinned = [torch.from_numpy(im).pin_memory() for im in images]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident detection (value) ----------------
    for i in range(max(W, 3 * NBUF)):  # every lane captures its CUDA graph on the second use of a buffer
        inst.detect_device(d_images[i % N_IMAGES].data_ptr(), w, h, i % NBUF)
    inst.wait_idle()
    counts = {}
    for i in range(N_IMAGES):
        inst.detect_device(d_images[i].data_ptr(), w, h, 0)
        counts[i] = inst.features_number(0)
    octaves = [inst.octave_resolution(o) for o in range(inst.nb_octaves())]
    ns = inst.config.nb_scales_per_octave

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = inst.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_feat = 0
    ev0.record(stream)
    for i in range(K):
        # a detection waits for the previous one of its lane (buffer index modulo the lane count); lanes overlap on the GPU
        inst.detect_device(d_images[i % N_IMAGES].data_ptr(), w, h, i % NBUF)
        n_feat += counts[i % N_IMAGES]
    inst.join_lanes()  # the instance stream waits (on the device) for every lane, so ev1 closes all K detections
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = inst.kernel_launch_count() - launches0

    # what the scale space costs per image in the pipelined schedule: the same timed loop with the blur launches left out
    # (vksiftx_setDebugSkip: analysis mode, the detections made while it is set are not used for anything else)
    def timed_loop(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3 * NBUF):
            inst.detect_device(d_images[i % N_IMAGES].data_ptr(), w, h, i % NBUF)
        inst.wait_idle()
        torch.cuda.synchronize()
        e0.record(stream)
        for i in range(n):
            inst.detect_device(d_images[i % N_IMAGES].data_ptr(), w, h, i % NBUF)
        inst.join_lanes()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    t_all = timed_loop(K)
    inst.set_debug_skip(8)
    t_no_pyr = timed_loop(K)
    inst.set_debug_skip(0)
    pyr_marginal_ms = max(t_all - t_no_pyr, 1e-6)

    # stage times: CUDA events of the library on its own stream, one synchronised detection at a time
    inst.set_profiling(True)
    stage_acc = {}
    KP = max(3, min(K, 20))
    for i in range(KP):
        inst.detect_device(d_images[i % N_IMAGES].data_ptr(), w, h, i % 2)
        for k, v in inst.stage_times_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v / KP
    # per-launch times of the scale-space stage (event pair around every launch; traced runs are not timed runs)
    inst.set_launch_trace(True)
    inst.set_serial_schedule(True)  # one stream: the event pair around a launch then times that kernel alone
    trace_acc = {}
    KT = 5
    for i in range(KT + 1):
        inst.detect_device(d_images[i % N_IMAGES].data_ptr(), w, h, i % 2)
        tr = inst.launch_trace()
        if i == 0:
            continue
        seen = {}
        for name, t0, t1 in tr:
            seen[name] = seen.get(name, 0) + 1
            key = "%s #%d" % (name, seen[name]) if name.startswith("fast o0 r4") else name
            trace_acc[key] = trace_acc.get(key, 0.0) + (t1 - t0) / KT
    inst.set_serial_schedule(False)
    inst.set_launch_trace(False)
    inst.set_profiling(False)

    # ---------------- end to end through the reference API (e2e) ----------------
    # The caller's loop is software-pipelined over the instance's feature buffers, the way the reference's two-buffer API
    # is meant to be used: image i+1.. are submitted (vksift_detectFeatures copies from page-locked host memory) before the
    # features of image i are fetched (vksift_getFeaturesNumber + vksift_downloadFeatures into page-locked host memory).
    # Every step still pays its own H2D and D2H inside the timed region.
    out_pinned = [torch.empty(20000 * api.FEATURE_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(api.FEATURE_DTYPE)
                  for _ in range(NBUF)]

    def e2e_loop(n):
        feats, nbytes = 0, 0
        for i in range(n + NBUF - 1):
            if i < n:
                inst.detect_raw(pinned[i % N_IMAGES].data_ptr(), w, h, i % NBUF)
            j = i - (NBUF - 1)
            if j >= 0:
                f = inst.download_features(j % NBUF, out=out_pinned[j % NBUF])
                feats += len(f)
                nbytes += f.nbytes + 4
        return feats, nbytes

    e2e_loop(3 * NBUF)
    barrier()
    t0 = time.perf_counter()
    e2e_feat, d2h = e2e_loop(K)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    # ---------------- small images (configs[2] pattern: 640x480, 8 images per GPU) ----------------
    # One 640x480 detection leaves a B200 mostly idle; an instance with 8 feature buffers runs 8 detection lanes, so the 8
    # images of a GPU's share overlap.  Reported, not the headline.
    from vulkansift_b200.synth import C1
    small = [blob_image(**dict(C1, seed=C1["seed"] + i)) for i in range(8)]
    d_small = [torch.from_numpy(im).cuda() for im in small]
    sh_, sw_ = small[0].shape
    small_res = {}
    for n_buf in (1, 8):
        sinst = api.Instance(gpu_device_index=local_rank, input_image_max_size=sw_ * sh_, max_nb_sift_per_buffer=20000, sift_buffer_count=n_buf)
        for rep in range(3):
            for i in range(8):
                sinst.detect_device(d_small[i].data_ptr(), sw_, sh_, i % n_buf)
        sinst.wait_idle()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 20
        for rep in range(reps):
            for i in range(8):
                sinst.detect_device(d_small[i].data_ptr(), sw_, sh_, i % n_buf)
        sinst.wait_idle()
        dt = time.perf_counter() - t0
        if n_buf == 8:
            small_res["features_per_8_images"] = sum(sinst.features_number(i) for i in range(8))
        small_res["images_per_s_%d_lane%s" % (sinst.lane_count(), "" if n_buf == 1 else "s")] = 8 * reps / dt
        sinst.close()

    # ---------------- matcher (configs[3]) ----------------
    da, db = random_descriptors(MATCH_N, 1234), random_descriptors(MATCH_N, 1235)
    minst = api.Instance(gpu_device_index=local_rank, max_nb_sift_per_buffer=MATCH_N, input_image_max_size=1024 * 1024)
    mstream = torch.cuda.ExternalStream(minst.stream, device=local_rank)
    fa = np.zeros(MATCH_N, api.FEATURE_DTYPE)
    fb = np.zeros(MATCH_N, api.FEATURE_DTYPE)
    fa["descriptor"], fb["descriptor"] = da, db
    minst.upload_features(fa, 0)
    minst.upload_features(fb, 1)
    minst.set_profiling(True)
    for _ in range(W):
        minst.match(0, 1)
    minst.wait_idle()
    mk = max(K, 20)
    m_acc = 0.0
    mev0, mev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    mev0.record(mstream)
    for _ in range(mk):
        minst.match(0, 1)
        m_acc += minst.stage_times_ms()["match_2nn"]
    mev1.record(mstream)
    barrier()
    match_ms = mev0.elapsed_time(mev1) / mk
    match_kernel_ms = m_acc / mk
    # e2e matcher: upload both descriptor sets, match, download the rows
    t0 = time.perf_counter()
    for _ in range(5):
        minst.upload_features(fa, 0)
        minst.upload_features(fb, 1)
        minst.match(0, 1)
        minst.download_matches()
    match_e2e_s = (time.perf_counter() - t0) / 5

    # ---------------- all-pairs cross-image matching (configs[4] pattern), N > 1 only ----------------
    allpairs = None
    if world > 1:
        from vulkansift_b200 import dist as vdist
        own = blob_image(**dict(C2, seed=C2["seed"] + 100 + rank))
        inst.detect(own, 0)
        n_own = inst.features_number(0)
        reps = 5
        ap_t, gather_t = [], []
        for rep in range(reps + 1):
            barrier()
            t0 = time.perf_counter()
            counts_g, blocks_g = vdist.gather_instance_descriptors(inst, 0, capacity=8191)  # 8192-row slots: 1 MB per rank
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            vdist.match_against_peers(inst, 0, 1, counts_g, blocks_g, rank, world, download=True)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            if rep > 0:  # first repetition warms NCCL up
                gather_t.append(t1 - t0)
                ap_t.append(t2 - t0)
        ap = torch.tensor([sum(ap_t) / reps, sum(gather_t) / reps], dtype=torch.float64, device="cuda")
        rows = torch.tensor([float(n_own * (world - 1))], dtype=torch.float64, device="cuda")
        dist.all_reduce(ap, op=dist.ReduceOp.MAX)
        dist.all_reduce(rows, op=dist.ReduceOp.SUM)
        allpairs = {"workload": "one 1920x1080 image per GPU, NCCL all-gather of descriptor blocks, every GPU matches its features "
                                "against each of the %d other blocks (ordered pairs: %d)" % (world - 1, world * (world - 1)),
                    "value": rows.item() / ap[0].item(), "unit": "matches/s", "ms_total": 1e3 * ap[0].item(),
                    "ms_gather": 1e3 * ap[1].item(), "matched_rows": rows.item()}

    # ---------------- reduce over ranks ----------------
    vals = torch.tensor([dev_ms, e2e_s, match_ms, match_kernel_ms, stage_acc.get("pyramid_dog", 0.0)], dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(n_feat), float(e2e_feat)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_ms, e2e_s, match_ms, match_kernel_ms, pyr_ms = vals.tolist()
    n_feat_all, e2e_feat_all = sums.tolist()

    if rank == 0:
        hbm, tf_burst, tf_sust, peak_src = measured_peaks()
        alg = algorithmic_bytes_pyramid(w, h, octaves, ns)
        ach = alg / (pyr_ms * 1e-3) / 1e9 if pyr_ms > 0 else 0.0
        flops = 2.0 * MATCH_N * MATCH_N * 128
        m_ach = flops / (match_kernel_ms * 1e-3) / 1e12 if match_kernel_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": n_feat_all / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(world),
            "features_per_image": [counts[i] for i in range(N_IMAGES)],
            "e2e": {"value": e2e_feat_all / e2e_s, "unit": UNIT, "h2d_bytes_per_step": w * h, "d2h_bytes_per_step": d2h // K,
                    "ms_per_step": 1e3 * e2e_s / K, "pipelined_over_buffers": NBUF},
            "detection_lanes": inst.lane_count(),
            "latency_ms_one_detection_alone": stage_acc.get("detect_total", 0.0),
            "gpu_launches": launches,
            "stage_ms": {k: v for k, v in stage_acc.items() if k.startswith(("pyramid", "extrema", "orient", "descr", "detect"))},
            "roofline": dominant_kernel_roofline(trace_acc, octaves, hbm, peak_src),
            "roofline_stage": {"bound": "hbm", "kernel": "whole pyramid+DoG stage (all blur launches of all octaves, CUDA events of the library)",
                               "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "algorithmic_bytes": alg, "stage_ms": pyr_ms,
                               "peak_source": peak_src},
            "roofline_stage_pipelined": {"bound": "hbm", "kernel": "whole pyramid+DoG stage, marginal cost per image with all lanes busy: ms per "
                                         "image of the timed loop minus the same loop with the blur launches left out "
                                         "(vksiftx_setDebugSkip)", "achieved": alg / (pyr_marginal_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                         "frac": alg / (pyr_marginal_ms * 1e-3) / 1e9 / hbm, "algorithmic_bytes": alg,
                                         "marginal_ms": pyr_marginal_ms, "ms_per_image_all_stages": t_all,
                                         "ms_per_image_without_scale_space": t_no_pyr, "peak_source": peak_src},
            "scale_space_launches_us": {k: round(v, 2) for k, v in trace_acc.items()},
            "match": {"metric": "2nn_matches_per_sec", "value": world * MATCH_N / (match_ms * 1e-3), "unit": "matches/s",
                      "workload": "configs[3]: 10000 x 10000 x 128-D u8, tcgen05 kind::i8 path", "ms_per_match_call": match_ms,
                      "kernel_ms": match_kernel_ms, "e2e_value": MATCH_N / match_e2e_s, "e2e_ms": 1e3 * match_e2e_s,
                      "roofline": {"bound": "tensor", "achieved": m_ach, "peak": tf_burst, "unit": "TFLOP/s", "frac": m_ach / tf_burst,
                                   "flops": flops, "peak_source": peak_src + " bf16 dense burst (i8 operands run at 2x this rate)"}},
            "clocks": clocks,
            "small_images": dict(small_res, workload="configs[2] pattern on one GPU: 8 x 640x480 (upsampled, default config), images resident in "
                                                     "HBM, wall clock over 20 rounds; N > 1: every rank does the same (weak scaling)"),
        }
        line["config"]["schedule"] = ("%d feature buffers = %d detection lanes (own scale space each, 0.5 GB): a detection waits for the "
                                      "previous one of its lane only, consecutive images overlap on the GPU" % (NBUF, inst.lane_count()))
        if allpairs:
            line["allpairs"] = allpairs
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_port_baseline(images, threads=os.cpu_count() or 1)
            line["cpu_baseline"] = cb
            line["match"]["cpu_baseline"] = cpu_match_baseline(threads=os.cpu_count() or 1)
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
