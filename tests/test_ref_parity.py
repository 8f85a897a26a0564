"""Pinning of the oracle against the REFERENCE's own code.

tests/golden/ref_vectors.npz was produced by tests/golden/make_golden.py from oracle/_ref: the reference's seven
GLSL shaders and three host functions, read in place from /root/reference and executed on the CPU
(oracle/build_ref.py, oracle/glsl_emu.h).  The oracle has to reproduce every vector: bit-exact for host tables,
DoG, keypoints (all float fields), orientations, descriptors and matches; 1e-6 absolute for the blur, whose
reference implementation samples through normalized texture coordinates (rounding of (x+0.5)/W).
Where the harness library is present (this container, or a box that received oracle/_ref) more configurations
are compared live."""
import os

import numpy as np
import pytest

from golden.make_golden import FIELDS, IMAGE, canon

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


def _bits(a):
    return a.view(np.uint32) if a.dtype.kind == "f" else a


def _same(a, b, fields):
    assert len(a) == len(b), (len(a), len(b))
    for n in fields:
        assert np.array_equal(_bits(np.ascontiguousarray(a[n])), _bits(np.ascontiguousarray(b[n]))), n


def test_host_tables_equal_reference_functions(gold, oracle_mod):
    cfgs = {"default": {}, "ns5": dict(nb_scales_per_octave=5), "plain": dict(use_interpolated_blur=0),
            "noup_s22": dict(use_input_upsampling=0, seed_scale_sigma=2.2)}
    for tag, kw in cfgs.items():
        ks, k = oracle_mod.Oracle(**kw).kernel_table()
        assert np.array_equal(ks, gold["ktab_%s_size" % tag])
        assert np.array_equal(k.view(np.uint32), gold["ktab_%s_k" % tag].view(np.uint32)), tag
    for w, h in ((640, 480), (1920, 1080), (333, 251), (67, 135), (1000, 37)):
        for up in (1, 0):
            o = oracle_mod.Oracle(use_input_upsampling=up, input_image_max_size=3840 * 2160)
            o.detect(np.zeros((h, w), np.uint8))
            got = np.array([o.octave_resolution(i) for i in range(o.nb_octaves)], np.uint32).reshape(-1, 2)
            exp = gold["oct_%dx%d_%s" % (w, h, "up" if up else "noup")].reshape(-1, 2)
            # the reference function is fed max_octaves = 32; the instance additionally caps at its own maximum
            assert np.array_equal(got, exp[:len(got)]) and len(got) == min(len(exp), o.max_octaves)
    for mx, n in ((100000, 7), (100000, 5), (300, 5), (1000, 3)):
        sizes = {7: (1080, 1920), 5: (480, 640), 3: (480, 640)}[n]
        kw = dict(max_nb_sift_per_buffer=mx)
        if n == 3:
            kw["nb_octaves"] = 3
        o = oracle_mod.Oracle(**kw)
        o.detect(np.zeros(sizes, np.uint8))
        assert o.nb_octaves == n
        assert np.array_equal(o.section_capacity(), gold["sec_%d_%d" % (mx, n)])


def test_shader_stages_equal_reference_shaders(gold, oracle_mod):
    from vulkansift_b200.synth import blob_image
    img = blob_image(**IMAGE)
    o = oracle_mod.Oracle(max_nb_orientation_per_keypoint=0)
    f = o.detect(img)
    prim, kept = o.primary_counts(), o.section_counts()[1]
    assert o.nb_octaves == 3 and int(kept.sum()) == len(f)
    for oc in range(o.nb_octaves):
        off = int(kept[:oc].sum())
        _same(canon(f[off:off + prim[oc]].copy(), FIELDS[:-1]), gold["extract_o%d" % oc], FIELDS[:-1])   # ExtractKeypoints.comp
        _same(canon(f[off:off + kept[oc]].copy()), gold["orient_o%d" % oc], FIELDS)                       # ComputeOrientation.comp
        assert np.array_equal(f[off:off + kept[oc]]["descriptor"], gold["desc_ubc_o%d" % oc])             # ComputeDescriptors.comp
    D = np.stack([o.dog(2, s) for s in range(5)])
    assert np.array_equal(D.view(np.uint32), gold["dog_o2"].view(np.uint32))                              # DifferenceOfGaussian.comp
    for s in (1, 3, 5):
        assert np.abs(o.gaussian(2, s) - gold["blur_o2_s%d" % s]).max() < 1e-6                             # GaussianBlurInterpolated.comp
    ov = oracle_mod.Oracle(max_nb_orientation_per_keypoint=0, use_vlfeat_format=1)
    fv = ov.detect(img)
    keptv = ov.section_counts()[1]
    for oc in range(ov.nb_octaves):
        off = int(keptv[:oc].sum())
        assert np.array_equal(fv[off:off + keptv[oc]]["descriptor"], gold["desc_vlfeat_o%d" % oc])


def test_matcher_equals_reference_shader(gold, oracle_mod):
    m = oracle_mod.match_descriptors(gold["match_a"], gold["match_b"], 1)      # Get2NearestNeighbors.comp
    assert m.tobytes() == gold["match_out"].tobytes()


def test_live_reference_harness_more_configurations(oracle_mod):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    from vulkansift_b200.synth import blob_image
    img = blob_image(200, 150, 90, seed=21)
    for kw, rkw in ((dict(nb_scales_per_octave=5), dict(ns=5)), (dict(use_input_upsampling=0), dict(upsample=False)),
                    (dict(use_interpolated_blur=0), dict(interp=False))):
        o = oracle_mod.Oracle(max_nb_orientation_per_keypoint=0, **kw)
        f = o.detect(img)
        ns = o.ns
        up = kw.get("use_input_upsampling", 1)
        ks, k = ref.kernel_table(**rkw)
        prim, kept = o.primary_counts(), o.section_counts()[1]
        assert len(f) > 10
        for oc in range(o.nb_octaves):
            G = np.stack([o.gaussian(oc, s) for s in range(ns + 3)])
            D = np.stack([o.dog(oc, s) for s in range(ns + 2)])
            off = int(kept[:oc].sum())
            sec = f[off:off + kept[oc]]
            fe, found = ref.extract(D, oc - up)
            assert found == prim[oc]
            _same(canon(fe, FIELDS[:-1]), canon(sec[:prim[oc]].copy(), FIELDS[:-1]), FIELDS[:-1])
            p0 = sec[:prim[oc]].copy()
            p0["orientation"] = 0
            _same(canon(ref.orientation(G, p0, max_ori=0)), canon(sec.copy()), FIELDS)
            assert np.array_equal(ref.descriptor(G, sec)["descriptor"], sec["descriptor"])
            assert np.array_equal(ref.dog(G).view(np.uint32), D.view(np.uint32))
            interp = kw.get("use_interpolated_blur", 1)
            for s in range(1, ns + 3):
                # normalized-coordinate rounding grows with the image width (W * 2^-24 texels)
                assert np.abs(ref.blur(G[s - 1], k[s], ks[s], bool(interp)) - G[s]).max() < 4e-6
    # default orientation cap (4): identical sets whenever a keypoint has at most 4 peaks
    o4 = oracle_mod.Oracle()
    f4 = o4.detect(img)
    prim, kept = o4.primary_counts(), o4.section_counts()[1]
    for oc in range(o4.nb_octaves):
        G = np.stack([o4.gaussian(oc, s) for s in range(6)])
        off = int(kept[:oc].sum())
        sec = f4[off:off + kept[oc]]
        p0 = sec[:prim[oc]].copy()
        p0["orientation"] = 0
        r4 = ref.orientation(G, p0, max_ori=4)
        assert len(r4) == len(sec)
