"""ctypes binding of libvulkansift.so, mirroring the reference C API one to one.

Every vksift_* entry point of include/vulkansift/vulkansift.h is exposed under
its own name through `lib`; `Instance` is a small convenience wrapper that
keeps the call order of the reference examples (src/examples/test_sift_detect.cpp,
test_sift_match.cpp): loadVulkan -> createInstance -> detectFeatures ->
getFeaturesNumber -> downloadFeatures -> matchFeatures -> downloadMatches.

Errors: the reference reports errors of void functions through
vksift_Config.on_error_callback_function (C++ callers throw from it).  A ctypes
callback cannot unwind through C frames, so the Python callback records the
code and the wrapper raises VksiftError right after the C call returns.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvulkansift.so")

VKSIFT_SUCCESS, VKSIFT_INVALID_INPUT_ERROR, VKSIFT_VULKAN_ERROR = 0, 1, 2
VKSIFT_NO_LOG, VKSIFT_LOG_ERROR, VKSIFT_LOG_WARNING, VKSIFT_LOG_INFO, VKSIFT_LOG_DEBUG = range(5)
VKSIFT_DESCRIPTOR_FORMAT_UBC, VKSIFT_DESCRIPTOR_FORMAT_VLFEAT = 0, 1
VKSIFT_PYRAMID_PRECISION_FLOAT32, VKSIFT_PYRAMID_PRECISION_FLOAT16 = 0, 1
NB_STAGES = 8
STAGE_NAMES = ("pyramid_dog", "extrema", "orientation", "descriptor", "detect_total", "match_prepare", "match_2nn",
               "match_total")

FEATURE_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("scale_x", "<f4"), ("scale_y", "<f4"), ("scale_idx", "<u4"),
                          ("octave_idx", "<i4"), ("sigma", "<f4"), ("orientation", "<f4"), ("intensity", "<f4"),
                          ("descriptor", "u1", (128,))])
MATCH_DTYPE = np.dtype([("idx_a", "<u4"), ("idx_b1", "<u4"), ("idx_b2", "<u4"), ("dist_a_b1", "<f4"),
                        ("dist_a_b2", "<f4")])

ERROR_CALLBACK = C.CFUNCTYPE(None, C.c_int)


class ExternalWindowInfo(C.Structure):
    _fields_ = [("context", C.c_void_p), ("window", C.c_void_p)]


class Config(C.Structure):
    """vksift_Config (include/vulkansift/vulkansift_types.h), 88 bytes."""
    _fields_ = [("input_image_max_size", C.c_uint32), ("sift_buffer_count", C.c_uint32),
                ("max_nb_sift_per_buffer", C.c_uint32), ("use_input_upsampling", C.c_bool), ("nb_octaves", C.c_uint8),
                ("nb_scales_per_octave", C.c_uint8), ("input_image_blur_level", C.c_float),
                ("seed_scale_sigma", C.c_float), ("intensity_threshold", C.c_float), ("edge_threshold", C.c_float),
                ("max_nb_orientation_per_keypoint", C.c_uint32), ("descriptor_format", C.c_int),
                ("gpu_device_index", C.c_int32), ("use_hardware_interpolated_blur", C.c_bool),
                ("pyramid_precision_mode", C.c_int), ("on_error_callback_function", ERROR_CALLBACK),
                ("use_gpu_debug_functions", C.c_bool), ("gpu_debug_external_window_info", ExternalWindowInfo)]


class VksiftError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        name = {1: "VKSIFT_INVALID_INPUT_ERROR", 2: "VKSIFT_VULKAN_ERROR"}.get(code, str(code))
        super().__init__("%s in %s" % (name, where))


def _load_library(path=LIB_PATH, analysis=False):
    if not os.path.exists(path):
        raise ImportError("%s is missing: build it with `python -m vulkansift_b200.build` (needs nvcc). "
                          "There is no CPU fallback." % path)
    lib = C.CDLL(path)
    I, P = C.c_void_p, C.POINTER
    u32, u8 = C.c_uint32, C.c_uint8
    sig = {
        "vksift_loadVulkan": (C.c_int, []),
        "vksift_unloadVulkan": (None, []),
        "vksift_getAvailableGPUs": (None, [P(u32), C.c_void_p]),
        "vksift_setLogLevel": (None, [C.c_int]),
        "vksift_createInstance": (C.c_int, [P(I), P(Config)]),
        "vksift_destroyInstance": (None, [P(I)]),
        "vksift_getDefaultConfig": (Config, []),
        "vksift_detectFeatures": (None, [I, C.c_void_p, u32, u32, u32]),
        "vksift_matchFeatures": (None, [I, u32, u32]),
        "vksift_getFeaturesNumber": (u32, [I, u32]),
        "vksift_downloadFeatures": (None, [I, C.c_void_p, u32]),
        "vksift_uploadFeatures": (None, [I, C.c_void_p, u32, u32]),
        "vksift_getMatchesNumber": (u32, [I]),
        "vksift_downloadMatches": (None, [I, C.c_void_p]),
        "vksift_isBufferAvailable": (C.c_bool, [I, u32]),
        "vksift_getScaleSpaceNbOctaves": (u8, [I]),
        "vksift_getScaleSpaceOctaveResolution": (None, [I, u8, P(u32), P(u32)]),
        "vksift_downloadScaleSpaceImage": (None, [I, u8, u8, C.c_void_p]),
        "vksift_downloadDoGImage": (None, [I, u8, u8, C.c_void_p]),
        "vksift_presentDebugFrame": (None, [I]),
        # extensions (include/vksift_b200_ext.h)
        "vksiftx_getVersionString": (C.c_char_p, []),
        "vksiftx_getDeviceIndex": (C.c_int32, [I]),
        "vksiftx_getStream": (C.c_void_p, [I]),
        "vksiftx_detectFeaturesDevice": (None, [I, C.c_void_p, u32, u32, u32]),
        "vksiftx_waitIdle": (None, [I]),
        "vksiftx_getLaneCount": (u32, [I]),
        "vksiftx_joinLanes": (None, [I]),
        "vksiftx_getBufferDeviceView": (None, [I, u32, P(u32), P(C.c_void_p), P(C.c_void_p)]),
        "vksiftx_uploadDescriptorsDevice": (None, [I, C.c_void_p, u32, u32]),
        "vksiftx_copyDescriptorsToDevice": (u32, [I, u32, C.c_void_p, u32]),
        "vksiftx_getMatchesDevice": (C.c_void_p, [I]),
        "vksiftx_matchFeaturesAgainstDevice": (None, [I, u32, C.c_void_p, u32]),
        "vksiftx_matchFeaturesAgainstBlocks": (None, [I, u32, C.c_void_p, u32, C.c_uint64, P(u32), u32]),
        "vksiftx_downloadMatchesBlocks": (None, [I, C.c_void_p, u32]),
        "vksiftx_exchangeCreate": (C.c_bool, [I, u32, u32, u32, C.c_void_p]),
        "vksiftx_exchangeConnect": (C.c_bool, [I, C.c_void_p]),
        "vksiftx_exchangeAllGather": (C.c_bool, [I, u32, P(u32), P(C.c_void_p), P(C.c_uint64)]),
        "vksiftx_exchangeMatchAllPeers": (C.c_bool, [I, u32, P(u32)]),
        "vksiftx_exchangeDestroy": (None, [I]),
        "vksiftx_setProfiling": (None, [I, C.c_bool]),
        "vksiftx_getStageTimesMs": (None, [I, P(C.c_float)]),
        "vksiftx_matchFeaturesCrossChecked": (C.c_uint32, [I, C.c_uint32, C.c_uint32, C.c_float, P(C.c_uint32), C.c_uint32]),
        "vksiftx_setLaunchTrace": (None, [I, C.c_bool]),
        "vksiftx_setSerialSchedule": (None, [I, C.c_bool]),
        "vksiftx_getLaunchTrace": (C.c_uint32, [I, C.c_void_p, P(C.c_float), P(C.c_float), C.c_uint32]),
        "vksiftx_getKernelLaunchCount": (C.c_uint64, [I]),
        "vksiftx_getEffectiveTaps": (None, [I, C.c_void_p, C.c_void_p]),
        "vksiftx_getSectionCapacities": (None, [I, u32, C.c_void_p]),
        "vksiftx_setMatcherImpl": (None, [I, C.c_int32]),
    }
    if analysis:  # libvulkansift_analysis.so only (-DVKS_ANALYSIS)
        sig["vksiftx_setDebugSkip"] = (None, [I, C.c_int32])
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTED_SYMBOLS = _load_library()

_loaded = False


def load():
    """vksift_loadVulkan(): raises VksiftError when no B200-class CUDA device is usable."""
    global _loaded
    if not _loaded:
        r = lib.vksift_loadVulkan()
        if r != VKSIFT_SUCCESS:
            raise VksiftError(r, "vksift_loadVulkan")
        _loaded = True


def unload():
    global _loaded
    lib.vksift_unloadVulkan()
    _loaded = False


def available_gpus():
    n = C.c_uint32(0)
    lib.vksift_getAvailableGPUs(C.byref(n), None)
    names = (C.c_char * 256 * max(1, n.value))()
    lib.vksift_getAvailableGPUs(C.byref(n), names)
    return [names[i].value.decode() for i in range(n.value)]


def default_config():
    return lib.vksift_getDefaultConfig()


class Instance:
    """One vksift_Instance (one GPU).  Keyword arguments override vksift_Config fields."""

    _lib = None  # the product library; vulkansift_b200.analysis.Instance binds the analysis build instead

    def __init__(self, **overrides):
        if self._lib is None:
            type(self)._lib = lib
        self._ensure_loaded()
        self._errors = []
        self._cb = ERROR_CALLBACK(lambda code: self._errors.append(int(code)))
        self.config = self._lib.vksift_getDefaultConfig()
        for k, v in overrides.items():
            if not hasattr(self.config, k):
                raise AttributeError("vksift_Config has no field %r" % k)
            setattr(self.config, k, v)
        self.config.on_error_callback_function = self._cb
        self._h = C.c_void_p(None)
        r = self._lib.vksift_createInstance(C.byref(self._h), C.byref(self.config))
        if r != VKSIFT_SUCCESS:
            raise VksiftError(r, "vksift_createInstance")

    # -- plumbing
    def _ensure_loaded(self):
        load()

    def _check(self, where):
        if self._errors:
            code = self._errors[0]
            self._errors.clear()
            raise VksiftError(code, where)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.vksift_destroyInstance(C.byref(self._h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- reference API
    def detect(self, image, buffer_id=0):
        """vksift_detectFeatures: image is a (h,w) uint8 numpy array (host memory)."""
        image = np.ascontiguousarray(image, np.uint8)
        h, w = image.shape
        self._lib.vksift_detectFeatures(self._h, image.ctypes.data, w, h, buffer_id)
        self._check("vksift_detectFeatures")

    def detect_raw(self, ptr, w, h, buffer_id=0):
        self._lib.vksift_detectFeatures(self._h, ptr, w, h, buffer_id)
        self._check("vksift_detectFeatures")

    def features_number(self, buffer_id=0):
        n = self._lib.vksift_getFeaturesNumber(self._h, buffer_id)
        self._check("vksift_getFeaturesNumber")
        return n

    def download_features(self, buffer_id=0, out=None):
        """vksift_getFeaturesNumber + vksift_downloadFeatures.  `out`: optional caller-owned FEATURE_DTYPE array (e.g. a view
        of page-locked memory) with room for the features; the filled prefix is returned."""
        n = self.features_number(buffer_id)
        if out is None:
            out = np.empty(n, FEATURE_DTYPE)
        else:
            assert out.dtype == FEATURE_DTYPE and len(out) >= n
            out = out[:n]
        self._lib.vksift_downloadFeatures(self._h, out.ctypes.data, buffer_id)
        self._check("vksift_downloadFeatures")
        return out

    def upload_features(self, feats, buffer_id=0):
        feats = np.ascontiguousarray(feats)
        assert feats.dtype == FEATURE_DTYPE
        self._lib.vksift_uploadFeatures(self._h, feats.ctypes.data, len(feats), buffer_id)
        self._check("vksift_uploadFeatures")

    def match(self, buffer_a=0, buffer_b=1):
        self._lib.vksift_matchFeatures(self._h, buffer_a, buffer_b)
        self._check("vksift_matchFeatures")

    def matches_number(self):
        return self._lib.vksift_getMatchesNumber(self._h)

    def download_matches(self):
        out = np.zeros(self.matches_number(), MATCH_DTYPE)
        self._lib.vksift_downloadMatches(self._h, out.ctypes.data)
        self._check("vksift_downloadMatches")
        return out

    def is_buffer_available(self, buffer_id=0):
        return bool(self._lib.vksift_isBufferAvailable(self._h, buffer_id))

    def nb_octaves(self):
        return self._lib.vksift_getScaleSpaceNbOctaves(self._h)

    def octave_resolution(self, octave):
        w, h = C.c_uint32(0), C.c_uint32(0)
        self._lib.vksift_getScaleSpaceOctaveResolution(self._h, octave, C.byref(w), C.byref(h))
        self._check("vksift_getScaleSpaceOctaveResolution")
        return w.value, h.value

    def _download_layer(self, fn, name, octave, scale):
        if 0 <= octave < self.nb_octaves():
            w, h = self.octave_resolution(octave)
        else:
            w, h = 1, 1
        out = np.zeros((h, w), np.float32)
        fn(self._h, octave, scale, out.ctypes.data)
        self._check(name)
        return out

    def download_scale_space_image(self, octave, scale):
        return self._download_layer(self._lib.vksift_downloadScaleSpaceImage, "vksift_downloadScaleSpaceImage", octave, scale)

    def download_dog_image(self, octave, scale):
        return self._download_layer(self._lib.vksift_downloadDoGImage, "vksift_downloadDoGImage", octave, scale)

    def present_debug_frame(self):
        self._lib.vksift_presentDebugFrame(self._h)

    # -- extensions
    @property
    def device_index(self):
        return self._lib.vksiftx_getDeviceIndex(self._h)

    @property
    def stream(self):
        return self._lib.vksiftx_getStream(self._h)

    def detect_device(self, dev_ptr, w, h, buffer_id=0):
        self._lib.vksiftx_detectFeaturesDevice(self._h, dev_ptr, w, h, buffer_id)
        self._check("vksiftx_detectFeaturesDevice")

    def wait_idle(self):
        self._lib.vksiftx_waitIdle(self._h)

    def lane_count(self):
        return int(self._lib.vksiftx_getLaneCount(self._h))

    def join_lanes(self):
        self._lib.vksiftx_joinLanes(self._h)

    def buffer_device_view(self, buffer_id=0):
        n, d, hd = C.c_uint32(0), C.c_void_p(None), C.c_void_p(None)
        self._lib.vksiftx_getBufferDeviceView(self._h, buffer_id, C.byref(n), C.byref(d), C.byref(hd))
        self._check("vksiftx_getBufferDeviceView")
        return n.value, d.value, hd.value

    def upload_descriptors_device(self, dev_ptr, n, buffer_id=0):
        self._lib.vksiftx_uploadDescriptorsDevice(self._h, dev_ptr, n, buffer_id)
        self._check("vksiftx_uploadDescriptorsDevice")

    def copy_descriptors_to_device(self, buffer_id, dev_ptr, capacity):
        n = self._lib.vksiftx_copyDescriptorsToDevice(self._h, buffer_id, dev_ptr, capacity)
        self._check("vksiftx_copyDescriptorsToDevice")
        return n

    def match_against_device(self, buffer_a, dev_ptr, n_b):
        """2-NN of buffer_a's features against n_b descriptors read in place from device memory."""
        self._lib.vksiftx_matchFeaturesAgainstDevice(self._h, buffer_a, dev_ptr, n_b)
        self._check("vksiftx_matchFeaturesAgainstDevice")

    def match_against_blocks(self, buffer_a, dev_ptr, counts, block_stride_bytes, skip_block=0xFFFFFFFF):
        """2-NN of buffer_a's features against every descriptor block (block j = counts[j] rows at dev_ptr + j * stride), enqueued
        back to back; fetch the results with download_matches_blocks()."""
        arr = (C.c_uint32 * len(counts))(*[int(c) for c in counts])
        self._n_blocks = len(counts)
        self._blocks_na = None
        self._lib.vksiftx_matchFeaturesAgainstBlocks(self._h, buffer_a, dev_ptr, len(counts), int(block_stride_bytes), arr, int(skip_block))
        self._check("vksiftx_matchFeaturesAgainstBlocks")

    def download_matches_blocks(self, n_rows, out=None):
        """(n_blocks, n_rows) MATCH_DTYPE array; n_rows = feature count of buffer A of the last match_against_blocks."""
        nb = self._n_blocks
        if out is None:
            out = np.zeros((nb, n_rows), MATCH_DTYPE)
        assert out.dtype == MATCH_DTYPE and out.size >= nb * n_rows
        self._lib.vksiftx_downloadMatchesBlocks(self._h, out.ctypes.data, nb)
        self._check("vksiftx_downloadMatchesBlocks")
        return out.reshape(-1)[:nb * n_rows].reshape(nb, n_rows)

    # -- descriptor exchange over NVLink peer memory (one process per GPU; see vulkansift_b200.dist.PeerExchange) --
    def exchange_create(self, rank, world, slot_rows):
        """Allocate this rank's receive region; returns its 64-byte CUDA IPC handle (bytes)."""
        h = (C.c_uint8 * 64)()
        ok = self._lib.vksiftx_exchangeCreate(self._h, rank, world, slot_rows, C.cast(h, C.c_void_p))
        self._check("vksiftx_exchangeCreate")
        assert ok
        self._xc_world = world
        return bytes(h)

    def exchange_connect(self, handles):
        """handles: world x 64 bytes in rank order."""
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        ok = self._lib.vksiftx_exchangeConnect(self._h, C.cast(buf, C.c_void_p))
        self._check("vksiftx_exchangeConnect")
        assert ok

    def exchange_allgather(self, buffer_id):
        """Collective.  Returns (counts, device pointer of block 0, block stride in bytes)."""
        counts = (C.c_uint32 * self._xc_world)()
        ptr, stride = C.c_void_p(), C.c_uint64()
        self._lib.vksiftx_exchangeAllGather(self._h, buffer_id, counts, C.byref(ptr), C.byref(stride))
        self._check("vksiftx_exchangeAllGather")
        return [int(c) for c in counts], ptr.value, int(stride.value)

    def exchange_match_all_peers(self, buffer_id):
        """Collective: exchange + searches against every peer block; fetch with download_matches_blocks().  Returns the counts."""
        counts = (C.c_uint32 * self._xc_world)()
        self._lib.vksiftx_exchangeMatchAllPeers(self._h, buffer_id, counts)
        self._check("vksiftx_exchangeMatchAllPeers")
        self._n_blocks = self._xc_world
        return [int(c) for c in counts]

    def exchange_destroy(self):
        self._lib.vksiftx_exchangeDestroy(self._h)

    def matches_device(self):
        return self._lib.vksiftx_getMatchesDevice(self._h)

    def set_profiling(self, enabled=True):
        self._lib.vksiftx_setProfiling(self._h, bool(enabled))

    def stage_times_ms(self):
        t = (C.c_float * NB_STAGES)()
        self._lib.vksiftx_getStageTimesMs(self._h, t)
        return dict(zip(STAGE_NAMES, [float(v) for v in t]))

    def match_cross_checked(self, buf_a, buf_b, lowe_ratio=0.75):
        """(n,2) uint32 pairs (idx in A, idx in B): mutual nearest neighbours passing the ratio test in both directions."""
        cap = max(1, self.features_number(buf_a))
        out = np.zeros((cap, 2), np.uint32)
        n = int(self._lib.vksiftx_matchFeaturesCrossChecked(self._h, buf_a, buf_b, lowe_ratio, out.ctypes.data_as(C.POINTER(C.c_uint32)), cap))
        self._check("vksiftx_matchFeaturesCrossChecked")
        return out[:min(n, cap)]

    def set_launch_trace(self, enabled=True):
        self._lib.vksiftx_setLaunchTrace(self._h, bool(enabled))

    def set_serial_schedule(self, enabled=True):
        self._lib.vksiftx_setSerialSchedule(self._h, bool(enabled))

    def launch_trace(self, capacity=256):
        """[(name, start_us, end_us)] of the scale-space launches of the last (traced) detection."""
        names = (C.c_char * 32 * capacity)()
        t0 = (C.c_float * capacity)()
        t1 = (C.c_float * capacity)()
        n = int(self._lib.vksiftx_getLaunchTrace(self._h, C.cast(names, C.c_void_p), t0, t1, capacity))
        return [(bytes(names[i]).split(b"\0")[0].decode(), float(t0[i]), float(t1[i])) for i in range(min(n, capacity))]

    def kernel_launch_count(self):
        return int(self._lib.vksiftx_getKernelLaunchCount(self._h))

    def effective_taps(self):
        n = self.config.nb_scales_per_octave + 3
        radius = np.zeros(n, np.uint32)
        taps = np.zeros((n, 21), np.float32)
        self._lib.vksiftx_getEffectiveTaps(self._h, radius.ctypes.data, taps.ctypes.data)
        return radius, taps

    def section_capacities(self, buffer_id=0):
        caps = np.zeros(16, np.uint32)
        self._lib.vksiftx_getSectionCapacities(self._h, buffer_id, caps.ctypes.data)
        self._check("vksiftx_getSectionCapacities")
        return caps

    def set_matcher_impl(self, impl):
        self._lib.vksiftx_setMatcherImpl(self._h, int(impl))
