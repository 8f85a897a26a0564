"""Drop-in boundary: the shared library loads on a CPU-only box, exports every symbol include/*.h declares,
and the public structs have the reference's binary layout.  No compute is called here."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def _declared_symbols():
    names = []
    for hdr in ("vulkansift/vulkansift.h", "vksift_b200_ext.h"):
        txt = open(os.path.join(INC, hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        txt = re.sub(r"#ifdef VKS_ANALYSIS.*?#endif", "", txt, flags=re.S)  # entry points of the analysis build only
        names += re.findall(r"VKSIFT_EXPORT[^;(]*?\b(vksiftx?_\w+)\s*\(", txt)
    return sorted(set(names))


def test_library_builds_loads_and_exports_everything():
    from vulkansift_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len([n for n in declared if n.startswith("vksift_")]) == 20  # the reference's 20 entry points
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    from vulkansift_b200 import api
    assert set(api.EXPORTED_SYMBOLS) == set(declared)


def test_product_library_has_no_stage_skipping_switches():
    """The ablation switch (vksiftx_setDebugSkip, stages left out, invalid results) only exists in the analysis build;
    the library that is benchmarked and shipped neither exports it nor reads any debug environment variable for it."""
    from vulkansift_b200 import build
    build.build()
    prod = ctypes.CDLL(build.LIB)
    ana = ctypes.CDLL(build.LIB_ANALYSIS)
    assert not hasattr(prod, "vksiftx_setDebugSkip")
    assert hasattr(ana, "vksiftx_setDebugSkip")
    blob = open(build.LIB, "rb").read()
    for needle in (b"VKSIFT_DEBUG_SKIP", b"VKSIFT_MEGA", b"setDebugSkip"):
        assert needle not in blob, needle


def test_struct_layout_matches_reference_abi():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "vulkansift/vulkansift.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(vksift_Feature), sizeof(vksift_Match_2NN), sizeof(vksift_Config), _Alignof(vksift_Config));
#define O(f) printf("%s %zu\n", #f, offsetof(vksift_Config, f));
  O(input_image_max_size) O(sift_buffer_count) O(max_nb_sift_per_buffer) O(use_input_upsampling) O(nb_octaves)
  O(nb_scales_per_octave) O(input_image_blur_level) O(seed_scale_sigma) O(intensity_threshold) O(edge_threshold)
  O(max_nb_orientation_per_keypoint) O(descriptor_format) O(gpu_device_index) O(use_hardware_interpolated_blur)
  O(pyramid_precision_mode) O(on_error_callback_function) O(use_gpu_debug_functions) O(gpu_debug_external_window_info)
  printf("desc %zu\n", offsetof(vksift_Feature, descriptor));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-std=c11", "-I", INC, c, "-o", exe], check=True)  # the header must be plain C
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split("\n")
    assert out[0].split() == ["164", "20", "88", "8"]
    offs = dict(l.split() for l in out[1:] if l)
    expected = {"input_image_max_size": 0, "sift_buffer_count": 4, "max_nb_sift_per_buffer": 8, "use_input_upsampling": 12,
                "nb_octaves": 13, "nb_scales_per_octave": 14, "input_image_blur_level": 16, "seed_scale_sigma": 20,
                "intensity_threshold": 24, "edge_threshold": 28, "max_nb_orientation_per_keypoint": 32, "descriptor_format": 36,
                "gpu_device_index": 40, "use_hardware_interpolated_blur": 44, "pyramid_precision_mode": 48,
                "on_error_callback_function": 56, "use_gpu_debug_functions": 64, "gpu_debug_external_window_info": 72, "desc": 36}
    assert {k: int(v) for k, v in offs.items()} == expected
    from vulkansift_b200 import api
    assert ctypes.sizeof(api.Config) == 88
    for k, v in expected.items():
        if k != "desc":
            assert getattr(api.Config, k).offset == v, k
    assert api.FEATURE_DTYPE.itemsize == 164 and api.MATCH_DTYPE.itemsize == 20


def test_default_config_values():
    from vulkansift_b200 import api
    c = api.default_config()  # vulkansift.c:47-64
    assert (c.input_image_max_size, c.sift_buffer_count, c.max_nb_sift_per_buffer) == (1920 * 1080, 2, 100000)
    assert (c.use_input_upsampling, c.nb_octaves, c.nb_scales_per_octave) == (True, 0, 3)
    assert abs(c.input_image_blur_level - 0.5) < 1e-7 and abs(c.seed_scale_sigma - 1.6) < 1e-6
    assert abs(c.intensity_threshold - 0.04) < 1e-7 and c.edge_threshold == 10.0
    assert c.max_nb_orientation_per_keypoint == 4 and c.descriptor_format == api.VKSIFT_DESCRIPTOR_FORMAT_UBC
    assert c.gpu_device_index == -1 and c.use_hardware_interpolated_blur and c.pyramid_precision_mode == 0
    assert not c.use_gpu_debug_functions


def test_no_gpu_fails_loudly_not_silently():
    """Without a CUDA device the product must refuse to work (no CPU fallback)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    code = ("from vulkansift_b200 import api\n"
            "api.lib.vksift_setLogLevel(0)\n"
            "try:\n api.load()\n print('LOADED')\nexcept api.VksiftError as e:\n print('ERR', e.code)\n"
            "try:\n api.Instance()\n print('CREATED')\nexcept api.VksiftError as e:\n print('ERR', e.code)\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True).stdout.split("\n")
    assert out[0] == "ERR 2" and out[1] == "ERR 2", out


def test_plain_c_caller_links_against_the_library(tmp_path):
    """A C program written against the reference's header links against the B200 library unchanged (no compute call)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "vulkansift_b200", "lib")
    for name in ("detect_match", "perf_runtime", "perf_matching", "exchange_pairs"):  # the README-style caller and the port of the reference's perf_sift_runtime driver
        out = tmp_path / name
        r = subprocess.run(["gcc", "-std=c11", os.path.join(root, "examples", name + ".c"), "-I" + os.path.join(root, "include"), "-L" + lib_dir,
                            "-lvulkansift", "-Wl,-rpath," + lib_dir, "-lm", "-o", str(out)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
