"""GPU parity of the 2-NN matcher (tcgen05 path and SIMT cross-check) vs the oracle.

Indices must be identical (integer work bit-exact); distances are sqrtf of exact
integers on both sides, so they are compared bit-exactly as well.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from vulkansift_b200 import api as a
    a.load()
    a.lib.vksift_setLogLevel(a.VKSIFT_LOG_WARNING)
    return a


def _feats(api, desc):
    f = np.zeros(len(desc), api.FEATURE_DTYPE)
    f["descriptor"] = desc
    f["x"] = np.arange(len(desc))
    return f


def _match(api, inst, da, db, impl):
    inst.set_matcher_impl(impl)
    inst.upload_features(_feats(api, da), 0)
    inst.upload_features(_feats(api, db), 1)
    inst.match(0, 1)
    assert inst.matches_number() == len(da)
    return inst.download_matches()


def _check(got, exp, what):
    for name in ("idx_a", "idx_b1", "idx_b2"):
        bad = np.flatnonzero(got[name] != exp[name])
        assert len(bad) == 0, "%s: %s differs on %d rows, first row %d got %d expected %d (d1 %g/%g d2 %g/%g)" % (
            what, name, len(bad), bad[0], got[name][bad[0]], exp[name][bad[0]], got["dist_a_b1"][bad[0]], exp["dist_a_b1"][bad[0]],
            got["dist_a_b2"][bad[0]], exp["dist_a_b2"][bad[0]])
    for name in ("dist_a_b1", "dist_a_b2"):
        assert np.array_equal(got[name].view(np.uint32), exp[name].view(np.uint32)), "%s: %s" % (what, name)


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("shape", [(1, 2), (5, 2), (100, 3), (128, 128), (129, 257), (1000, 777), (3000, 3100)])
def test_random_descriptors(api, oracle_mod, shape, impl):
    from vulkansift_b200.synth import random_descriptors
    na, nb = shape
    da, db = random_descriptors(na, 100 + na), random_descriptors(nb, 200 + nb)
    inst = api.Instance(max_nb_sift_per_buffer=4096)
    got = _match(api, inst, da, db, impl)
    _check(got, oracle_mod.match_descriptors(da, db), "%s impl %d" % (shape, impl))
    inst.close()


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
def test_ties_follow_the_shader_rule(api, oracle_mod, impl):
    """Duplicated B rows force exact ties, including between b=0 and b=1 (Get2NearestNeighbors.comp:69-96)."""
    rng = np.random.default_rng(7)
    base = rng.integers(0, 256, (40, 128), dtype=np.uint8)
    db = np.concatenate([base[:1], base[:1], base, base[::-1], base[:5]])  # b0 == b1, many duplicates
    da = np.concatenate([base[:20], rng.integers(0, 256, (300, 128), dtype=np.uint8), base[:1]])
    inst = api.Instance(max_nb_sift_per_buffer=1024)
    got = _match(api, inst, da, db, impl)
    _check(got, oracle_mod.match_descriptors(da, db), "ties impl %d" % impl)
    inst.close()


def test_extreme_values(api, oracle_mod):
    da = np.concatenate([np.zeros((3, 128), np.uint8), np.full((3, 128), 255, np.uint8)])
    db = np.concatenate([np.full((130, 128), 255, np.uint8), np.zeros((130, 128), np.uint8)])
    db[5, 3] = 254
    inst = api.Instance(max_nb_sift_per_buffer=1024)
    for impl in (1, 0):
        _check(_match(api, inst, da, db, impl), oracle_mod.match_descriptors(da, db), "extreme impl %d" % impl)
    inst.close()


def test_empty_a_and_too_small_b(api):
    inst = api.Instance(max_nb_sift_per_buffer=1024)
    inst.upload_features(np.zeros(0, api.FEATURE_DTYPE), 0)
    inst.upload_features(_feats(api, np.zeros((4, 128), np.uint8)), 1)
    inst.match(0, 1)
    assert inst.matches_number() == 0 and len(inst.download_matches()) == 0
    inst.upload_features(_feats(api, np.zeros((4, 128), np.uint8)), 0)
    inst.upload_features(_feats(api, np.zeros((1, 128), np.uint8)), 1)
    with pytest.raises(api.VksiftError) as e:
        inst.match(0, 1)
    assert e.value.code == api.VKSIFT_INVALID_INPUT_ERROR
    inst.close()


def test_detected_features_match_and_ratio_test(api, oracle_mod, c1_image):
    """The reference example's flow (test_sift_match.cpp:67-107): detect two images, match both ways,
    cross-check + Lowe ratio 0.75 on true L2 distances."""
    shifted = np.roll(np.roll(c1_image, 7, axis=1), 5, axis=0)
    inst = api.Instance()
    inst.detect(c1_image, 0)
    inst.detect(shifted, 1)
    fa, fb = inst.download_features(0), inst.download_features(1)
    inst.match(0, 1)
    m01 = inst.download_matches()
    inst.match(1, 0)
    m10 = inst.download_matches()
    _check(m01, oracle_mod.match_features(fa, fb), "detected 0->1")
    _check(m10, oracle_mod.match_features(fb, fa), "detected 1->0")
    good = [(m["idx_a"], m["idx_b1"]) for m in m01
            if m10["idx_b1"][m["idx_b1"]] == m["idx_a"] and m["dist_a_b1"] / m["dist_a_b2"] < 0.75]
    assert len(good) > 300
    dx = np.array([fb["x"][b] - fa["x"][a] for a, b in good])
    dy = np.array([fb["y"][b] - fa["y"][a] for a, b in good])
    inl = (np.abs(dx - 7) < 1.0) & (np.abs(dy - 5) < 1.0)
    assert inl.mean() > 0.9
    inst.close()


def test_c4_full_size_10k_x_10k(api, oracle_mod):
    """BASELINE configs[3].  Full-size: tcgen05 vs SIMT on every row, and vs the oracle on a row sample;
    size-independent property: d1 <= d2 and both are true minima of a brute-force numpy check on 64 rows."""
    from vulkansift_b200.synth import random_descriptors
    da, db = random_descriptors(10000, 1234), random_descriptors(10000, 1235)
    inst = api.Instance(max_nb_sift_per_buffer=10000)
    tc = _match(api, inst, da, db, 0)
    simt = _match(api, inst, da, db, 1)
    _check(tc, simt, "C4 tcgen05 vs simt")
    sel = np.arange(0, 10000, 19)
    exp = oracle_mod.match_descriptors(da[sel], db)
    sub = tc[sel].copy()
    sub["idx_a"] = np.arange(len(sel))
    _check(sub, exp, "C4 sample vs oracle")
    assert np.all(tc["dist_a_b1"] <= tc["dist_a_b2"])
    rows = sel[:64]
    d2 = ((da[rows, None, :].astype(np.int32) - db[None, :, :].astype(np.int32)) ** 2).sum(-1)
    assert np.array_equal(np.sort(d2, axis=1)[:, 0], np.round(tc["dist_a_b1"][rows].astype(np.float64) ** 2).astype(np.int64))
    inst.close()


def test_cross_checked_matcher_equals_the_reference_callers_loop(api):
    """vksiftx_matchFeaturesCrossChecked == the CPU loop of src/examples/test_sift_match.cpp:90-107 applied to the two
    downloaded 2-NN lists (mutual nearest neighbour, Lowe ratio 0.75 in both directions), pairs in increasing idx_a."""
    from vulkansift_b200.synth import blob_image
    img = blob_image(800, 600, 500, seed=31)
    shifted = np.roll(img, (3, 5), axis=(0, 1))
    with api.Instance() as inst:
        inst.detect(img, 0)
        inst.detect(shifted, 1)
        inst.match(0, 1)
        m12 = inst.download_matches()
        inst.match(1, 0)
        m21 = inst.download_matches()
        exp = []
        old = np.seterr(all="ignore")  # 0/0 for identical descriptors: NaN < 0.75 is False, as in the C++ loop
        for i in range(len(m12)):
            j = int(m12[i]["idx_b1"])
            if int(m21[j]["idx_b1"]) != i:
                continue
            if not (np.float32(m12[i]["dist_a_b1"]) / np.float32(m12[i]["dist_a_b2"]) < 0.75):
                continue
            if not (np.float32(m21[j]["dist_a_b1"]) / np.float32(m21[j]["dist_a_b2"]) < 0.75):
                continue
            exp.append((i, j))
        np.seterr(**old)
        got = inst.match_cross_checked(0, 1, 0.75)
        assert len(exp) > 50
        assert [tuple(int(v) for v in r) for r in got] == exp
        # the retained match list is the A->B one
        assert inst.download_matches().tobytes() == m12.tobytes()


@pytest.mark.parametrize("shape", [(700, 2), (1000, 777), (3000, 3100)])
def test_match_against_device_descriptors_in_place(api, oracle_mod, shape):
    """vksiftx_matchFeaturesAgainstDevice: B read in place from caller-owned device memory (a peer's block of the NCCL
    all-gather) gives the records vksift_matchFeatures gives for the same B in a feature buffer."""
    import torch
    from vulkansift_b200.synth import random_descriptors
    na, nb = shape
    da, db = random_descriptors(na, 300 + na), random_descriptors(nb, 400 + nb)
    inst = api.Instance(max_nb_sift_per_buffer=4096)
    inst.upload_features(_feats(api, da), 0)
    slot = torch.zeros((2, 4096, 128), dtype=torch.uint8, device="cuda")  # B sits in the second slot, exactly nb rows used
    slot[1, :nb] = torch.from_numpy(db).cuda()
    slot[1, nb:] = 255  # rows past nb must never be read as candidates
    torch.cuda.synchronize()
    inst.match_against_device(0, slot[1].data_ptr(), nb)
    got = inst.download_matches()
    _check(got, oracle_mod.match_descriptors(da, db), "in place %s" % (shape,))
    with pytest.raises(api.VksiftError) as e:
        inst.match_against_device(0, slot[1].data_ptr() + 64, nb)
    assert e.value.code == api.VKSIFT_INVALID_INPUT_ERROR
    with pytest.raises(api.VksiftError) as e:
        inst.match_against_device(0, slot[1].data_ptr(), 1)
    assert e.value.code == api.VKSIFT_INVALID_INPUT_ERROR
    inst.close()


def test_match_against_blocks_batched_equals_oracle(api, oracle_mod):
    """vksiftx_matchFeaturesAgainstBlocks (the all-pairs step: A against every gathered peer block, enqueued back to back, one
    download): every list equals the oracle's, the skipped block and a block with one row come back zero-filled."""
    import torch
    from vulkansift_b200.synth import random_descriptors
    cap = 1023
    counts = [700, 1, 1023, 333]
    da = random_descriptors(900, 51)
    blocks = torch.zeros((len(counts), cap + 1, 128), dtype=torch.uint8, device="cuda")
    host = []
    for j, n in enumerate(counts):
        d = random_descriptors(n, 60 + j)
        if j == 2:
            d[5] = d[4]          # a tie inside a block
            d[:2] = da[10]       # b = 0 and b = 1 tie for A row 10
        host.append(d)
        blocks[j, :n] = torch.from_numpy(d).cuda()
    torch.cuda.synchronize()
    with api.Instance(max_nb_sift_per_buffer=4096) as inst:
        inst.upload_features(_feats(api, da), 0)
        inst.match_against_blocks(0, blocks.data_ptr(), counts, blocks.stride(0), skip_block=3)
        res = inst.download_matches_blocks(len(da))
        assert res.shape == (4, 900)
        _check(res[0], oracle_mod.match_descriptors(da, host[0]), "block 0")
        _check(res[2], oracle_mod.match_descriptors(da, host[2]), "block 2")
        assert not res[1].view(np.uint8).any() and not res[3].view(np.uint8).any()
        # a second call with other blocks skipped reuses the buffers
        inst.match_against_blocks(0, blocks.data_ptr(), counts, blocks.stride(0), skip_block=0)
        res = inst.download_matches_blocks(len(da))
        _check(res[3], oracle_mod.match_descriptors(da, host[3]), "block 3")
        assert not res[0].view(np.uint8).any()
        # the ordinary single-result API is untouched
        inst.upload_features(_feats(api, host[0]), 1)
        inst.match(0, 1)
        _check(inst.download_matches(), oracle_mod.match_descriptors(da, host[0]), "plain match afterwards")


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
def test_large_distances_follow_the_float_comparison(api, oracle_mod, impl):
    """Beyond d^2 = 2^22 distinct squared distances round to the same float sqrt and the shader's strict '<' on floats keeps the
    earlier scan position (Get2NearestNeighbors.comp:82-96).  Byte descriptors far apart (not SIFT descriptors: |a - b| >= 2048)
    with squared distances that differ by one in that range must come out like the oracle, which compares floats."""
    rng = np.random.default_rng(9)
    na, nb = 40, 300
    # d^2 = 73 * 255^2 + 68^2 + (v_a - v_b)^2 with v in {0, 1}: 4751449 and 4751450 have the same float square root (2179.782)
    da = np.zeros((na, 128), np.uint8)
    da[:, 127] = rng.integers(0, 2, na)
    db = np.zeros((nb, 128), np.uint8)
    db[:, :73] = 255
    db[:, 73] = 68
    db[:, 127] = rng.integers(0, 2, nb)
    assert np.sqrt(np.float32(4751449)) == np.sqrt(np.float32(4751450))
    exp = oracle_mod.match_descriptors(da, db)
    sq = exp["dist_a_b2"].astype(np.float64) ** 2
    assert sq.min() >= 2 ** 22
    with api.Instance(max_nb_sift_per_buffer=1024) as inst:
        got = _match(api, inst, da, db, impl)
    _check(got, exp, "large distances impl %d" % impl)


def test_back_to_back_searches_are_stream_ordered(api, oracle_mod):
    """vksift_matchFeatures no longer waits for the previous search on the host: searches issued back to back overlap on the GPU
    (the next MMA kernel starts under the previous merge).  The retained result is the last search's, an upload into a buffer a
    pending search reads waits for it, and every intermediate result is still correct when fetched in between."""
    from vulkansift_b200.synth import random_descriptors
    d = [random_descriptors(n, 300 + i) for i, n in enumerate((1500, 2100, 900, 2600))]
    exp = {(a, b): oracle_mod.match_descriptors(d[a], d[b]) for a, b in ((0, 1), (1, 0), (2, 3), (3, 2), (0, 3))}
    with api.Instance(max_nb_sift_per_buffer=4096, sift_buffer_count=4, input_image_max_size=1 << 20) as inst:
        for i in range(4):
            inst.upload_features(_feats(api, d[i]), i)
        for rep in range(3):  # norms cached from the second repetition on: every search may overlap its predecessor
            for a, b in ((0, 1), (1, 0), (2, 3), (3, 2), (0, 3)):
                inst.match(a, b)
            assert inst.matches_number() == len(d[0])
            _check(inst.download_matches(), exp[(0, 3)], "last of five searches, repetition %d" % rep)
        # results fetched in between
        for a, b in ((2, 3), (0, 1), (3, 2)):
            inst.match(a, b)
            _check(inst.download_matches(), exp[(a, b)], "search (%d, %d)" % (a, b))
        # an upload into B right behind the search: the search still sees the old descriptors
        inst.match(0, 1)
        inst.match(1, 0)
        inst.upload_features(_feats(api, d[2]), 0)
        _check(inst.download_matches(), exp[(1, 0)], "search pending during an upload")
        inst.match(0, 3)
        _check(inst.download_matches(), oracle_mod.match_descriptors(d[2], d[3]), "search after the upload")
