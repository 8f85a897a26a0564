"""ctypes binding of oracle/sift_oracle.c (test infrastructure, see sift_oracle.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

FEATURE_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("scale_x", "<f4"), ("scale_y", "<f4"), ("scale_idx", "<u4"),
                          ("octave_idx", "<i4"), ("sigma", "<f4"), ("orientation", "<f4"), ("intensity", "<f4"),
                          ("descriptor", "u1", (128,))])
MATCH_DTYPE = np.dtype([("idx_a", "<u4"), ("idx_b1", "<u4"), ("idx_b2", "<u4"), ("dist_a_b1", "<f4"),
                        ("dist_a_b2", "<f4")])
assert FEATURE_DTYPE.itemsize == 164 and MATCH_DTYPE.itemsize == 20


class OracleConfig(C.Structure):
    _fields_ = [("input_image_max_size", C.c_uint32), ("max_nb_sift_per_buffer", C.c_uint32),
                ("use_input_upsampling", C.c_int32), ("nb_octaves", C.c_int32), ("nb_scales_per_octave", C.c_int32),
                ("input_image_blur_level", C.c_float), ("seed_scale_sigma", C.c_float),
                ("intensity_threshold", C.c_float), ("edge_threshold", C.c_float),
                ("max_nb_orientation_per_keypoint", C.c_uint32), ("use_vlfeat_format", C.c_int32),
                ("use_interpolated_blur", C.c_int32), ("use_fp16_pyramid", C.c_int32), ("nb_threads", C.c_int32)]


def lib_path():
    return os.path.join(_HERE, "_build", "libsift_oracle.so")


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile).  Building the checker is not using it."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("sift_oracle.c", "sift_oracle.h", "Makefile"))
    src_m = max(src_m, os.path.getmtime(os.path.join(_HERE, "..", "include", "vksift_arith.h")))
    if force or not os.path.exists(lib_path()) or os.path.getmtime(lib_path()) < src_m:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return lib_path()


_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    lib = C.CDLL(build())
    P = C.POINTER
    lib.vkso_default_config.argtypes = [P(OracleConfig)]
    lib.vkso_create.argtypes = [P(OracleConfig)]
    lib.vkso_create.restype = C.c_void_p
    lib.vkso_destroy.argtypes = [C.c_void_p]
    lib.vkso_max_octaves.argtypes = [C.c_void_p]
    lib.vkso_kernel_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vkso_effective_taps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vkso_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    lib.vkso_detect.restype = C.c_uint32
    lib.vkso_nb_octaves.argtypes = [C.c_void_p]
    lib.vkso_octave_resolution.argtypes = [C.c_void_p, C.c_int, P(C.c_uint32), P(C.c_uint32)]
    for f in ("vkso_section_capacity", "vkso_primary_counts"):
        getattr(lib, f).argtypes = [C.c_void_p, C.c_void_p]
    lib.vkso_section_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vkso_get_features.argtypes = [C.c_void_p, C.c_void_p]
    for f in ("vkso_gaussian_layer", "vkso_dog_layer"):
        getattr(lib, f).argtypes = [C.c_void_p, C.c_int, C.c_int]
        getattr(lib, f).restype = P(C.c_float)
    lib.vkso_stage_seconds.argtypes = [C.c_void_p, C.c_void_p]
    lib.vkso_match.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
    lib.vkso_match_features.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
    lib.vkso_seed_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    lib.vkso_downsample_nearest.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    lib.vkso_expf.argtypes = [C.c_float]
    lib.vkso_expf.restype = C.c_float
    lib.vkso_exp2f.argtypes = [C.c_float]
    lib.vkso_exp2f.restype = C.c_float
    lib.vkso_atan2f.argtypes = [C.c_float, C.c_float]
    lib.vkso_atan2f.restype = C.c_float
    lib.vkso_sincosf.argtypes = [C.c_float, P(C.c_float), P(C.c_float)]
    lib.vkso_ceil_log2.argtypes = [C.c_float]
    lib.vkso_mirror.argtypes = [C.c_int, C.c_int]
    _lib = lib
    return lib


class _Arith:
    """Scalar probes of include/vksift_arith.h as compiled by gcc."""

    def expf(self, x):
        return _load().vkso_expf(float(x))

    def exp2f(self, x):
        return _load().vkso_exp2f(float(x))

    def atan2f(self, y, x):
        return _load().vkso_atan2f(float(y), float(x))

    def sincosf(self, t):
        s, c = C.c_float(), C.c_float()
        _load().vkso_sincosf(float(t), C.byref(s), C.byref(c))
        return s.value, c.value

    def ceil_log2(self, m):
        return _load().vkso_ceil_log2(float(m))

    def mirror(self, i, n):
        return _load().vkso_mirror(int(i), int(n))


arith = _Arith()


class Oracle:
    """One oracle context = one vksift_Config."""

    def __init__(self, **overrides):
        lib = _load()
        self.cfg = OracleConfig()
        lib.vkso_default_config(C.byref(self.cfg))
        for k, v in overrides.items():
            if not hasattr(self.cfg, k):
                raise AttributeError(k)
            setattr(self.cfg, k, v)
        self._ctx = lib.vkso_create(C.byref(self.cfg))
        self.ns = self.cfg.nb_scales_per_octave
        self.nb_features = 0

    def close(self):
        if self._ctx:
            _load().vkso_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # host-side tables
    @property
    def max_octaves(self):
        return _load().vkso_max_octaves(self._ctx)

    def kernel_table(self):
        n = self.ns + 3
        ksize = np.zeros(n, np.uint32)
        k = np.zeros((n, 20), np.float32)
        _load().vkso_kernel_table(self._ctx, ksize.ctypes.data, k.ctypes.data)
        return ksize, k

    def effective_taps(self):
        n = self.ns + 3
        radius = np.zeros(n, np.uint32)
        e = np.zeros((n, 21), np.float32)
        _load().vkso_effective_taps(self._ctx, radius.ctypes.data, e.ctypes.data)
        return radius, e

    # detection
    def detect(self, image):
        image = np.ascontiguousarray(image, np.uint8)
        assert image.ndim == 2
        h, w = image.shape
        self.nb_features = _load().vkso_detect(self._ctx, image.ctypes.data, w, h)
        return self.features()

    def features(self):
        out = np.zeros(self.nb_features, FEATURE_DTYPE)
        if self.nb_features:
            _load().vkso_get_features(self._ctx, out.ctypes.data)
        return out

    @property
    def nb_octaves(self):
        return _load().vkso_nb_octaves(self._ctx)

    def octave_resolution(self, o):
        w, h = C.c_uint32(), C.c_uint32()
        _load().vkso_octave_resolution(self._ctx, o, C.byref(w), C.byref(h))
        return w.value, h.value

    def section_capacity(self):
        caps = np.zeros(self.nb_octaves, np.uint32)
        _load().vkso_section_capacity(self._ctx, caps.ctypes.data)
        return caps

    def section_counts(self):
        found = np.zeros(self.nb_octaves, np.uint32)
        kept = np.zeros(self.nb_octaves, np.uint32)
        _load().vkso_section_counts(self._ctx, found.ctypes.data, kept.ctypes.data)
        return found, kept

    def primary_counts(self):
        p = np.zeros(self.nb_octaves, np.uint32)
        _load().vkso_primary_counts(self._ctx, p.ctypes.data)
        return p

    def _layer(self, fn, o, s):
        w, h = self.octave_resolution(o)
        ptr = fn(self._ctx, o, s)
        return np.ctypeslib.as_array(ptr, shape=(h, w)).copy()

    def gaussian(self, o, s):
        return self._layer(_load().vkso_gaussian_layer, o, s)

    def dog(self, o, s):
        return self._layer(_load().vkso_dog_layer, o, s)

    def stage_seconds(self):
        t = np.zeros(4, np.float64)
        _load().vkso_stage_seconds(self._ctx, t.ctypes.data)
        return dict(zip(("pyramid_dog", "extrema", "orientation", "descriptor"), t.tolist()))


def match_descriptors(a, b, nb_threads=0):
    """2-NN of every row of a (na,128 u8) in b (nb,128 u8), nb >= 2."""
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    assert a.shape[1] == 128 and b.shape[1] == 128 and b.shape[0] >= 2
    out = np.zeros(a.shape[0], MATCH_DTYPE)
    _load().vkso_match(a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], out.ctypes.data, nb_threads)
    return out


def match_features(fa, fb, nb_threads=0):
    fa = np.ascontiguousarray(fa)
    fb = np.ascontiguousarray(fb)
    assert fa.dtype == FEATURE_DTYPE and fb.dtype == FEATURE_DTYPE and len(fb) >= 2
    out = np.zeros(len(fa), MATCH_DTYPE)
    _load().vkso_match_features(fa.ctypes.data, len(fa), fb.ctypes.data, len(fb), out.ctypes.data, nb_threads)
    return out


def seed_image(image, dw, dh):
    """u8 UNORM upload + LINEAR blit to (dh, dw) (or the 1:1 conversion): the Vulkan fixed-function step in front of the first blur."""
    image = np.ascontiguousarray(image, np.uint8)
    out = np.zeros((dh, dw), np.float32)
    _load().vkso_seed_image(image.ctypes.data, image.shape[1], image.shape[0], out.ctypes.data, dw, dh)
    return out


def downsample_nearest(layer, dw, dh):
    """NEAREST blit of a Gaussian layer into the next octave's layer 0."""
    layer = np.ascontiguousarray(layer, np.float32)
    out = np.zeros((dh, dw), np.float32)
    _load().vkso_downsample_nearest(layer.ctypes.data, layer.shape[1], layer.shape[0], out.ctypes.data, dw, dh)
    return out
