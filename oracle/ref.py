"""ctypes binding of oracle/_ref/libvksift_ref.so (the reference's own shaders and host functions on the CPU).

TEST INFRASTRUCTURE ONLY.  `available()` is true when the library has been built (this container, where
/root/reference exists, or a GPU box that received the prebuilt oracle/_ref/)."""
import ctypes as C
import os

import numpy as np

from . import build_ref
from .oracle import FEATURE_DTYPE, MATCH_DTYPE

_lib = None


def available():
    return os.path.exists(build_ref.OUT) or build_ref.available()


def _load():
    global _lib
    if _lib is None:
        if build_ref.available():
            build_ref.build()
        _lib = C.CDLL(build_ref.OUT)
    return _lib


def kernel_table(ns=3, upsample=True, input_blur=0.5, seed_sigma=1.6, interp=True):
    ksize = np.zeros(ns + 3, np.uint32)
    k = np.zeros((ns + 3, 20), np.float32)
    _load().ref_host_kernel_table(C.c_uint32(ns), C.c_int(int(upsample)), C.c_float(input_blur), C.c_float(seed_sigma), C.c_int(int(interp)),
                                  C.c_void_p(ksize.ctypes.data), C.c_void_p(k.ctypes.data))
    return ksize, k


def octaves(w, h, upsample=True, max_octaves=32):
    ow, oh = np.zeros(64, np.uint32), np.zeros(64, np.uint32)
    f = _load().ref_host_octaves
    f.restype = C.c_uint32
    n = f(C.c_uint32(w), C.c_uint32(h), C.c_int(int(upsample)), C.c_uint32(max_octaves), C.c_void_p(ow.ctypes.data), C.c_void_p(oh.ctypes.data))
    return [(int(ow[i]), int(oh[i])) for i in range(n)]


def sections(max_feats, n_oct):
    caps = np.zeros(n_oct, np.uint32)
    _load().ref_host_sections(C.c_uint32(max_feats), C.c_uint32(n_oct), C.c_void_p(caps.ctypes.data))
    return caps


def blur_pass(img, kernel20, ksize, vertical, interp):
    img = np.ascontiguousarray(img, np.float32)
    out = np.zeros_like(img)
    k = np.ascontiguousarray(kernel20, np.float32)
    fn = _load().ref_blur_interp if interp else _load().ref_blur_plain
    fn(C.c_void_p(img.ctypes.data), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_void_p(out.ctypes.data), C.c_void_p(k.ctypes.data),
       C.c_uint(int(ksize)), C.c_int(int(vertical)))
    return out


def blur(img, kernel20, ksize, interp):
    """horizontal then vertical pass, as recorded by sift_detector.c:955-1001"""
    return blur_pass(blur_pass(img, kernel20, ksize, 0, interp), kernel20, ksize, 1, interp)


def dog(gauss_layers):
    g = np.ascontiguousarray(gauss_layers, np.float32)
    nl, h, w = g.shape
    out = np.zeros((nl - 1, h, w), np.float32)
    _load().ref_dog(C.c_void_p(g.ctypes.data), C.c_int(w), C.c_int(h), C.c_int(nl), C.c_void_p(out.ctypes.data))
    return out


def extract(dog_layers, octave_idx, seed_sigma=1.6, intensity_threshold=0.04, edge_threshold=10.0, max_feat=100000):
    d = np.ascontiguousarray(dog_layers, np.float32)
    nl, h, w = d.shape
    ns = nl - 2
    out = np.zeros(max_feat, FEATURE_DTYPE)
    f = _load().ref_extract
    f.restype = C.c_uint
    thr = np.float32(intensity_threshold) / np.float32(ns)  # sift_detector.c:1136
    found = f(C.c_void_p(d.ctypes.data), C.c_int(w), C.c_int(h), C.c_int(ns), C.c_int(octave_idx), C.c_float(seed_sigma), C.c_float(thr),
              C.c_float(edge_threshold), C.c_uint(max_feat), C.c_void_p(out.ctypes.data))
    return out[:min(found, max_feat)].copy(), found


def orientation(gauss_layers, feats, max_ori=4, max_feat=None):
    g = np.ascontiguousarray(gauss_layers, np.float32)
    nl, h, w = g.shape
    n = len(feats)
    max_feat = max_feat or (n * 36 + 1)
    buf = np.zeros(max_feat, FEATURE_DTYPE)
    buf[:n] = feats
    f = _load().ref_orientation
    f.restype = C.c_uint
    found = f(C.c_void_p(g.ctypes.data), C.c_int(w), C.c_int(h), C.c_int(nl), C.c_void_p(buf.ctypes.data), C.c_uint(n), C.c_uint(max_feat),
              C.c_uint(max_ori))
    return buf[:min(found, max_feat)].copy()


def descriptor(gauss_layers, feats, vlfeat=False):
    g = np.ascontiguousarray(gauss_layers, np.float32)
    nl, h, w = g.shape
    buf = np.ascontiguousarray(feats).copy()
    _load().ref_descriptor(C.c_void_p(g.ctypes.data), C.c_int(w), C.c_int(h), C.c_int(nl), C.c_void_p(buf.ctypes.data), C.c_uint(len(buf)),
                           C.c_uint(int(vlfeat)))
    return buf


def match(fa, fb):
    fa, fb = np.ascontiguousarray(fa), np.ascontiguousarray(fb)
    out = np.zeros(len(fa), MATCH_DTYPE)
    _load().ref_match(C.c_void_p(fa.ctypes.data), C.c_uint(len(fa)), C.c_void_p(fb.ctypes.data), C.c_uint(len(fb)), C.c_void_p(out.ctypes.data))
    return out
