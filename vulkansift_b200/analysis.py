"""Binding of the ANALYSIS build of the library (vulkansift_b200/lib/libvulkansift_analysis.so, -DVKS_ANALYSIS).

Same kernels and entry points as the product library plus vksiftx_setDebugSkip, which leaves stages of a detection out
so that tools/ablation.py and bench.py can measure what a stage costs in the pipelined schedule.  Detections made while a
bit is set return invalid results; nothing that is timed as a headline number or compared with the oracle runs here.
"""
import os

from . import api

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libvulkansift_analysis.so")
lib, EXPORTED_SYMBOLS = api._load_library(LIB_PATH, analysis=True)
_loaded = False


class Instance(api.Instance):
    _lib = lib

    def _ensure_loaded(self):
        global _loaded
        if not _loaded:  # the analysis library has its own process-global "API loaded" flag
            r = lib.vksift_loadVulkan()
            if r != api.VKSIFT_SUCCESS:
                raise api.VksiftError(r, "vksift_loadVulkan")
            lib.vksift_setLogLevel(api.VKSIFT_LOG_WARNING)
            _loaded = True

    def set_debug_skip(self, mask):
        """Bits: 1 descriptors, 2 orientation, 4 extrema + refinement + ordering, 8 scale space.  Results INVALID while set."""
        lib.vksiftx_setDebugSkip(self._h, int(mask))
