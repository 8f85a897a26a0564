"""GPU parity of the detection path: CUDA library (through the C ABI) vs the CPU oracle.

Integer / index work must be bit-exact; because CUDA kernels and oracle share the
fp32 arithmetic contract of include/vksift_arith.h, float fields are compared
bit-exactly too (tolerance 0, stricter than BASELINE.json's 1e-4 relative).
"""
import numpy as np
import pytest

from conftest import assert_features_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from vulkansift_b200 import api as a
    a.load()
    a.lib.vksift_setLogLevel(a.VKSIFT_LOG_WARNING)
    return a


def _oracle_cfg_from(inst_kwargs):
    m = {"use_input_upsampling": "use_input_upsampling", "nb_octaves": "nb_octaves",
         "nb_scales_per_octave": "nb_scales_per_octave", "input_image_blur_level": "input_image_blur_level",
         "seed_scale_sigma": "seed_scale_sigma", "intensity_threshold": "intensity_threshold",
         "edge_threshold": "edge_threshold", "max_nb_orientation_per_keypoint": "max_nb_orientation_per_keypoint",
         "input_image_max_size": "input_image_max_size", "max_nb_sift_per_buffer": "max_nb_sift_per_buffer"}
    out = {}
    for k, v in inst_kwargs.items():
        if k in m:
            out[m[k]] = int(v) if isinstance(v, bool) else v
        elif k == "descriptor_format":
            out["use_vlfeat_format"] = int(v)
        elif k == "use_hardware_interpolated_blur":
            out["use_interpolated_blur"] = int(v)
    return out


def _run_pair(api, oracle_mod, image, **kw):
    inst = api.Instance(**kw)
    orc = oracle_mod.Oracle(**_oracle_cfg_from(kw))
    exp = orc.detect(image)
    inst.detect(image, 0)
    got = inst.download_features(0)
    return inst, orc, got, exp


def test_host_tables_match_oracle(api, oracle_mod):
    for kw in ({}, {"nb_scales_per_octave": 5}, {"use_hardware_interpolated_blur": False},
               {"use_input_upsampling": False, "seed_scale_sigma": 2.2}):
        inst = api.Instance(**kw)
        orc = oracle_mod.Oracle(**_oracle_cfg_from(kw))
        r, t = inst.effective_taps()
        ro, to = orc.effective_taps()
        assert np.array_equal(r, ro)
        assert np.array_equal(t.view(np.uint32), to.view(np.uint32))
        inst.close()


def test_pyramid_and_dog_bit_exact_c1(api, oracle_mod, c1_image):
    inst, orc, got, exp = _run_pair(api, oracle_mod, c1_image)
    assert inst.nb_octaves() == orc.nb_octaves == 5
    ns = 3
    for o in range(orc.nb_octaves):
        assert inst.octave_resolution(o) == orc.octave_resolution(o)
        for s in range(ns + 3):
            g = inst.download_scale_space_image(o, s)
            e = orc.gaussian(o, s)
            bad = np.count_nonzero(g.view(np.uint32) != e.view(np.uint32))
            assert bad == 0, "Gaussian octave %d scale %d: %d px differ, max abs %g" % (o, s, bad, np.abs(g - e).max())
        for s in range(ns + 2):
            g = inst.download_dog_image(o, s)
            e = orc.dog(o, s)
            assert np.array_equal(g.view(np.uint32), e.view(np.uint32)), "DoG octave %d scale %d" % (o, s)
    inst.close()


def test_features_bit_exact_c1(api, oracle_mod, c1_image):
    inst, orc, got, exp = _run_pair(api, oracle_mod, c1_image)
    assert len(exp) == 601  # calibrated count of the C1 workload
    assert_features_equal(got, exp, "C1")
    caps = inst.section_capacities(0)[:orc.nb_octaves]
    assert np.array_equal(caps, orc.section_capacity())
    inst.close()


@pytest.mark.parametrize("kw", [
    {"use_input_upsampling": False},
    {"nb_scales_per_octave": 5},
    {"descriptor_format": 1},
    {"use_hardware_interpolated_blur": False},
    {"max_nb_orientation_per_keypoint": 0},
    {"max_nb_orientation_per_keypoint": 1},
    {"nb_octaves": 2},
    {"intensity_threshold": 0.01, "edge_threshold": 5.0},
], ids=lambda kw: ",".join("%s=%s" % kv for kv in kw.items()))
def test_features_bit_exact_config_variants(api, oracle_mod, kw):
    from vulkansift_b200.synth import blob_image
    img = blob_image(320, 240, 150, seed=3)
    inst, orc, got, exp = _run_pair(api, oracle_mod, img, **kw)
    assert len(exp) > 20
    assert_features_equal(got, exp, str(kw))
    inst.close()


@pytest.mark.parametrize("wh", [(333, 251), (67, 135), (1000, 37), (32, 32), (129, 257)])
def test_ragged_sizes(api, oracle_mod, wh):
    from vulkansift_b200.synth import blob_image
    w, h = wh
    img = blob_image(w, h, max(4, w * h // 600), seed=11)
    inst, orc, got, exp = _run_pair(api, oracle_mod, img)
    assert inst.nb_octaves() == orc.nb_octaves
    for o in range(orc.nb_octaves):
        assert inst.octave_resolution(o) == orc.octave_resolution(o)
        for s in (0, 3, 5):
            assert np.array_equal(inst.download_scale_space_image(o, s).view(np.uint32), orc.gaussian(o, s).view(np.uint32)), (o, s)
    assert_features_equal(got, exp, str(wh))
    inst.close()


def test_constant_and_noise_images(api, oracle_mod):
    inst = api.Instance()
    inst.detect(np.full((480, 640), 128, np.uint8), 0)
    assert inst.features_number(0) == 0
    assert len(inst.download_features(0)) == 0
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (240, 320), dtype=np.uint8)
    orc = oracle_mod.Oracle()
    exp = orc.detect(img)
    inst.detect(img, 1)
    assert_features_equal(inst.download_features(1), exp, "noise")
    inst.close()


def test_section_overflow_is_deterministic(api, oracle_mod, c1_image):
    kw = {"max_nb_sift_per_buffer": 300}
    inst, orc, got, exp = _run_pair(api, oracle_mod, c1_image, **kw)
    found, kept = orc.section_counts()
    assert (found > kept).any(), "workload must overflow at least one section"
    assert_features_equal(got, exp, "overflow")
    inst.close()


def test_resolution_change_and_two_buffers(api, oracle_mod, c1_image):
    from vulkansift_b200.synth import blob_image
    small = blob_image(200, 160, 60, seed=9)
    inst = api.Instance()
    orc = oracle_mod.Oracle()
    inst.detect(c1_image, 0)
    inst.detect(small, 1)
    inst.detect(c1_image, 0)
    assert_features_equal(inst.download_features(1), orc.detect(small), "small")
    assert_features_equal(inst.download_features(0), orc.detect(c1_image), "c1 again")
    inst.close()


def test_device_resident_input_equals_host_input(api, c1_image):
    import torch
    inst = api.Instance()
    inst.detect(c1_image, 0)
    a = inst.download_features(0)
    d = torch.from_numpy(c1_image).cuda()
    torch.cuda.synchronize()
    inst.detect_device(d.data_ptr(), c1_image.shape[1], c1_image.shape[0], 1)
    b = inst.download_features(1)
    assert_features_equal(b, a, "device input")
    inst.close()


def test_upload_download_round_trip(api, c1_image):
    inst = api.Instance()
    inst.detect(c1_image, 0)
    f = inst.download_features(0)
    inst.upload_features(f, 1)
    assert inst.features_number(1) == len(f)
    g = inst.download_features(1)
    assert f.tobytes() == g.tobytes()
    inst.close()


def test_error_semantics(api, c1_image):
    inst = api.Instance(sift_buffer_count=2, input_image_max_size=640 * 480)
    with pytest.raises(api.VksiftError) as e:
        inst.detect(c1_image, 2)  # any idx >= NB_BUFF must fail (test_sift_error_handling.cpp:56-57)
    assert e.value.code == api.VKSIFT_INVALID_INPUT_ERROR
    with pytest.raises(api.VksiftError):
        inst.detect(np.zeros((1000, 1000), np.uint8), 0)  # larger than input_image_max_size
    with pytest.raises(api.VksiftError):
        inst.detect(np.zeros((16, 16), np.uint8), 0)  # < 1024 pixels
    with pytest.raises(api.VksiftError):
        inst.download_scale_space_image(0, 6)
    with pytest.raises(api.VksiftError):
        inst.download_dog_image(0, 5)
    with pytest.raises(api.VksiftError):
        inst.download_scale_space_image(99, 0)
    # the instance stays usable after INVALID_INPUT errors
    inst.detect(c1_image, 0)
    assert inst.features_number(0) == 601
    inst.close()
    # invalid configuration is reported by the return value
    with pytest.raises(api.VksiftError) as e:
        api.Instance(seed_scale_sigma=0.5)  # below 2x input blur with upsampling
    assert e.value.code == api.VKSIFT_INVALID_INPUT_ERROR


def test_c2_full_size_matches_oracle(api, oracle_mod):
    """BASELINE configs[1]: 1920x1080, upsampling, sigma0 1.6 -- full parity at full size."""
    from vulkansift_b200.synth import blob_image, C2
    img = blob_image(**C2)
    inst, orc, got, exp = _run_pair(api, oracle_mod, img)
    assert inst.nb_octaves() == 7
    assert 2500 <= len(exp) <= 3600, len(exp)
    assert_features_equal(got, exp, "C2")
    g = inst.download_dog_image(0, 2)
    assert np.array_equal(g.view(np.uint32), orc.dog(0, 2).view(np.uint32))
    inst.close()


@pytest.mark.parametrize("env", ["VKSIFT_GRAPH=1", "VKSIFT_GRAPH=0", "VKSIFT_NO_SPLIT=1", "VKSIFT_NO_PDL=1", "VKSIFT_STRIP=1"])
def test_alternative_schedules_are_bit_exact(api, oracle_mod, env, monkeypatch):
    """The scale space can be scheduled several ways (CUDA-graph replay (the default with several lanes) or eager launches,
    no stage overlap, no programmatic dependent launch, per-layer launches instead of the multi-layer strip kernel): same
    bytes out."""
    from vulkansift_b200.synth import blob_image
    monkeypatch.setenv(*env.split("="))
    imgs = [blob_image(640, 480, 400, seed=11), blob_image(1000, 700, 600, seed=12)]
    for img in imgs:
        with api.Instance() as inst:
            orc = oracle_mod.Oracle()
            exp = orc.detect(img)
            for rep in range(3):  # the graph is captured on the second use of a buffer and replayed on the third
                inst.detect(img, 0)
                got = inst.download_features(0)
                assert_features_equal(got, exp)


def test_launch_trace_reports_every_scale_space_launch(api):
    from vulkansift_b200.synth import blob_image
    img = blob_image(640, 480, 300, seed=5)
    with api.Instance() as inst:
        inst.set_launch_trace(True)
        inst.detect(img, 0)
        tr = inst.launch_trace()
        inst.set_launch_trace(False)
    assert len(tr) >= 5
    assert all(t1 >= t0 >= 0.0 for _, t0, t1 in tr)
    assert any(name.startswith(("fast o0", "strip o0")) for name, _, _ in tr)


def test_fp16_pyramid_precision_mode_bit_exact(api, oracle_mod):
    """VKSIFT_PYRAMID_PRECISION_FLOAT16 (sift_memory.c:139): every Gaussian and DoG value goes through binary16 when it is
    stored (SURVEY B-D11); layers, keypoints and descriptors must equal the oracle's fp16 mode bit for bit."""
    from vulkansift_b200.synth import blob_image
    for img in (blob_image(640, 480, 400, seed=21), blob_image(1920, 1080, 1200, seed=22)):
        with api.Instance(pyramid_precision_mode=api.VKSIFT_PYRAMID_PRECISION_FLOAT16) as inst:
            orc = oracle_mod.Oracle(use_fp16_pyramid=1)
            exp = orc.detect(img)
            inst.detect(img, 0)
            got = inst.download_features(0)
            assert len(exp) > 50
            assert_features_equal(got, exp)
            for o in range(inst.nb_octaves()):
                for s in (0, 3, 5):
                    g = inst.download_scale_space_image(o, s)
                    e = orc.gaussian(o, s)
                    assert np.array_equal(g.view(np.uint32), e.view(np.uint32)), (o, s)
                d = inst.download_dog_image(o, 2)
                e = orc.dog(o, 2)
                assert np.array_equal(d.view(np.uint32), e.view(np.uint32)), o


def test_serial_schedule_is_bit_exact(api, oracle_mod):
    from vulkansift_b200.synth import blob_image
    img = blob_image(1000, 700, 600, seed=12)
    with api.Instance() as inst:
        exp = oracle_mod.Oracle().detect(img)
        inst.set_serial_schedule(True)
        inst.detect(img, 0)
        assert_features_equal(inst.download_features(0), exp)
        inst.set_serial_schedule(False)
        inst.detect(img, 1)
        assert_features_equal(inst.download_features(1), exp)


def test_detection_lanes_overlap_and_stay_bit_exact(api, oracle_mod):
    """Detections into different buffers run on different lanes (own scale space, scratch, streams) without waiting for
    each other; every result must equal the oracle's, whatever the interleaving, resolution mix and input memory kind."""
    import torch
    from vulkansift_b200.synth import blob_image
    shapes = [(320, 240), (257, 193), (320, 240), (400, 300)]
    imgs = [blob_image(w, h, 120 + 10 * i, seed=20 + i) for i, (w, h) in enumerate(shapes)]
    orc = oracle_mod.Oracle()
    exp = [orc.detect(im) for im in imgs]
    inst = api.Instance(sift_buffer_count=4)
    assert inst.lane_count() == 4
    pinned = [torch.from_numpy(im).pin_memory() for im in imgs]
    dev = [torch.from_numpy(im).cuda() for im in imgs]
    torch.cuda.synchronize()
    for rep in range(3):
        # all four enqueued back to back: pageable host, page-locked host (read by the copy engine) and device inputs
        for b, im in enumerate(imgs):
            h, w = im.shape
            if (b + rep) % 3 == 0:
                inst.detect(im, b)
            elif (b + rep) % 3 == 1:
                inst.detect_raw(pinned[b].data_ptr(), w, h, b)
            else:
                inst.detect_device(dev[b].data_ptr(), w, h, b)
        # the scale space shown is the one of the most recent detection
        assert inst.octave_resolution(0) == (2 * shapes[3][0], 2 * shapes[3][1])
        for b in (2, 0, 3, 1):
            assert_features_equal(inst.download_features(b), exp[b], "rep %d buffer %d" % (rep, b))
        # a lane is reused by a different image while the others still hold theirs
        inst.detect(imgs[1], 0)
        assert_features_equal(inst.download_features(0), exp[1], "reused lane")
        assert_features_equal(inst.download_features(1), exp[1], "untouched buffer")
    # matching waits for every lane
    inst.detect(imgs[0], 0)
    inst.detect(imgs[2], 2)
    inst.match(0, 2)
    m = inst.download_matches()
    assert m.tobytes() == oracle_mod.match_features(exp[0], exp[2]).tobytes()
    inst.close()


def test_single_lane_override(api, oracle_mod, c1_image, monkeypatch):
    monkeypatch.setenv("VKSIFT_LANES", "1")
    inst = api.Instance()
    assert inst.lane_count() == 1
    inst.detect(c1_image, 0)
    inst.detect(c1_image[::-1].copy(), 1)
    a, b = inst.download_features(0), inst.download_features(1)
    inst.close()
    monkeypatch.delenv("VKSIFT_LANES")
    inst = api.Instance()
    assert inst.lane_count() == 2
    inst.detect(c1_image, 0)
    inst.detect(c1_image[::-1].copy(), 1)
    assert_features_equal(inst.download_features(0), a, "lanes vs single lane, buffer 0")
    assert_features_equal(inst.download_features(1), b, "lanes vs single lane, buffer 1")
    inst.close()


def test_benchmarked_configuration_c2_eight_lanes_graph_replay(api, oracle_mod):
    """The configuration bench.py times: 1920x1080 (C2), 8 feature buffers = 8 detection lanes, device-resident inputs, every lane
    replaying its CUDA graph (captured on the second use of a buffer).  Three rounds of 8 detections in flight, every buffer
    compared with the oracle after every round."""
    import torch
    from vulkansift_b200.synth import blob_image, C2
    imgs = [blob_image(**dict(C2, seed=C2["seed"] + i)) for i in range(4)]
    orc = oracle_mod.Oracle()
    exp = [orc.detect(im) for im in imgs]
    assert all(2500 <= len(e) <= 3800 for e in exp)
    h, w = imgs[0].shape
    dev = [torch.from_numpy(im).cuda() for im in imgs]
    torch.cuda.synchronize()
    with api.Instance(input_image_max_size=w * h, sift_buffer_count=8) as inst:
        assert inst.lane_count() == 8
        for rep in range(3):
            for b in range(8):
                inst.detect_device(dev[(b + rep) % 4].data_ptr(), w, h, b)
            for b in range(8):
                assert_features_equal(inst.download_features(b), exp[(b + rep) % 4], "round %d buffer %d" % (rep, b))


def test_small_images_eight_lanes(api, oracle_mod):
    """configs[2] pattern on one GPU: 8 x 640x480 in flight on 8 lanes, three rounds (graph replay from the second)."""
    from vulkansift_b200.synth import blob_image, C1
    imgs = [blob_image(**dict(C1, seed=C1["seed"] + i)) for i in range(8)]
    orc = oracle_mod.Oracle()
    exp = [orc.detect(im) for im in imgs]
    with api.Instance(input_image_max_size=640 * 480, sift_buffer_count=8) as inst:
        for rep in range(3):
            for b in range(8):
                inst.detect(imgs[(b + 3 * rep) % 8], b)
            for b in range(8):
                assert_features_equal(inst.download_features(b), exp[(b + 3 * rep) % 8], "round %d buffer %d" % (rep, b))


@pytest.mark.parametrize("queue", [None, "16"], ids=["default_queue", "overflowing_queue"])
@pytest.mark.parametrize("max_feats", [50, 300])
def test_many_raw_extrema_small_buffer_keeps_lowest_accepted_keys(api, oracle_mod, max_feats, queue, monkeypatch):
    """A blocky noise image yields tens of thousands of strict extrema and ~5.6 k accepted keypoints in octave 0 while the buffer holds 50
    or 300: the kept set must be the oracle's (lowest (s, y, x) keys among the ACCEPTED keypoints of each octave, then the
    section clamp), whatever order the GPU found them in.  (Round 1 queued raw extrema in a max_nb_sift_per_buffer-sized
    buffer and dropped the overflow in arrival order.)  With VKSIFT_RAW_QUEUE=16 every large octave overflows its extrema
    queue (256 entries minimum per octave) and takes the queue-less slow path; the result must not change."""
    if queue:
        monkeypatch.setenv("VKSIFT_RAW_QUEUE", queue)
    rng = np.random.default_rng(17)
    img = np.kron(rng.integers(0, 256, (100, 133), dtype=np.uint8), np.ones((3, 3), np.uint8))  # 300x399 field of 3x3 blocks
    kw = {"max_nb_sift_per_buffer": max_feats}
    orc = oracle_mod.Oracle(**kw)
    exp = orc.detect(img)
    found, kept = orc.section_counts()
    assert found[0] > 4 * kept[0], "octave 0 must overflow by a wide margin"
    with api.Instance(**kw) as inst:
        for rep in range(3):
            inst.detect(img, rep % 2)
            assert_features_equal(inst.download_features(rep % 2), exp, "max %d rep %d" % (max_feats, rep))


def test_extrema_queue_overflow_path_equals_default_on_c1(api, oracle_mod, c1_image, monkeypatch):
    """The slow path that needs no extrema queue (taken when an octave yields more strict extrema than its queue share) on an
    ordinary image: same features, same order, same bytes as the oracle."""
    monkeypatch.setenv("VKSIFT_RAW_QUEUE", "16")
    with api.Instance() as inst:
        exp = oracle_mod.Oracle().detect(c1_image)
        for rep in range(2):
            inst.detect(c1_image, rep)
            assert_features_equal(inst.download_features(rep), exp, "rep %d" % rep)


def test_too_many_scales_is_rejected_at_creation(api):
    """nb_scales_per_octave beyond what the extrema scan can stage must fail in vksift_createInstance with an input error,
    not later inside every vksift_detectFeatures."""
    with pytest.raises(api.VksiftError) as e:
        api.Instance(nb_scales_per_octave=12)
    assert e.value.code == api.VKSIFT_INVALID_INPUT_ERROR
    with api.Instance(nb_scales_per_octave=8) as inst:
        inst.detect(np.full((100, 100), 7, np.uint8), 0)
        assert inst.features_number(0) == 0
