"""Generate tests/golden/ref_vectors.npz from the REFERENCE's own code run on the CPU (oracle/_ref, built by
oracle/build_ref.py from /root/reference).  Run in the build container:  python tests/golden/make_golden.py

Stage inputs are the oracle's intermediate images for a seeded synthetic picture; stage outputs are what the
reference's shaders / host functions compute from them.  tests/test_ref_parity.py then requires the oracle to
reproduce every stored output (bit-exact, except the blur where the reference samples through normalized
texture coordinates: 1e-6 absolute)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import ref  # noqa: E402
from vulkansift_b200.synth import blob_image  # noqa: E402

IMAGE = dict(width=160, height=120, n_blobs=60, seed=5)
FIELDS = ("scale_x", "scale_y", "scale_idx", "octave_idx", "sigma", "intensity", "x", "y", "orientation")


def canon(a, fields=FIELDS):
    keys = [a[f].view(np.uint32) if a[f].dtype.kind == "f" else a[f] for f in fields]
    return a[np.lexsort(keys[::-1])]


def main():
    out = {}
    # host tables (setupGaussianKernels, updateScaleSpaceInfo, updateBufferInfo)
    for tag, kw in (("default", {}), ("ns5", dict(ns=5)), ("plain", dict(interp=False)), ("noup_s22", dict(upsample=False, seed_sigma=2.2))):
        ks, k = ref.kernel_table(**kw)
        out["ktab_%s_size" % tag], out["ktab_%s_k" % tag] = ks, k
    for w, h in ((640, 480), (1920, 1080), (333, 251), (67, 135), (1000, 37)):
        out["oct_%dx%d_up" % (w, h)] = np.array(ref.octaves(w, h, True), np.uint32)
        out["oct_%dx%d_noup" % (w, h)] = np.array(ref.octaves(w, h, False), np.uint32)
    for mx, n in ((100000, 7), (100000, 5), (300, 5), (1000, 3)):
        out["sec_%d_%d" % (mx, n)] = ref.sections(mx, n)

    img = blob_image(**IMAGE)
    for tag, okw, vl in (("ubc", dict(max_nb_orientation_per_keypoint=0), False),
                         ("vlfeat", dict(max_nb_orientation_per_keypoint=0, use_vlfeat_format=1), True)):
        o = oracle.Oracle(**okw)
        f = o.detect(img)
        prim, kept = o.primary_counts(), o.section_counts()[1]
        ks, k = ref.kernel_table()
        for oc in range(o.nb_octaves):
            G = np.stack([o.gaussian(oc, s) for s in range(6)])
            D = np.stack([o.dog(oc, s) for s in range(5)])
            off = int(kept[:oc].sum())
            sec = f[off:off + kept[oc]]
            if tag == "ubc":
                fe, found = ref.extract(D, oc - 1)
                out["extract_o%d" % oc] = canon(fe, FIELDS[:-1])
                p0 = f[off:off + prim[oc]].copy()
                p0["orientation"] = 0
                p0["descriptor"] = 0
                out["orient_o%d" % oc] = canon(ref.orientation(G, p0, max_ori=0))
                if oc == 2:
                    out["dog_o2"] = ref.dog(G)
                    for s in (1, 3, 5):
                        out["blur_o2_s%d" % s] = ref.blur(G[s - 1], k[s], ks[s], True)
            out["desc_%s_o%d" % (tag, oc)] = ref.descriptor(G, sec, vlfeat=vl)["descriptor"]
    rng = np.random.default_rng(1)
    fa, fb = np.zeros(50, oracle.FEATURE_DTYPE), np.zeros(70, oracle.FEATURE_DTYPE)
    fa["descriptor"], fb["descriptor"] = rng.integers(0, 256, (50, 128)), rng.integers(0, 256, (70, 128))
    fb["descriptor"][1] = fb["descriptor"][0]
    fa["descriptor"][3] = fb["descriptor"][0]
    fb["descriptor"][40] = fb["descriptor"][20]
    out["match_a"], out["match_b"], out["match_out"] = fa["descriptor"], fb["descriptor"], ref.match(fa, fb)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
