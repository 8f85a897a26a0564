"""Build oracle/_ref/libvksift_ref.so: the REFERENCE's own code for the hot path, executed on the CPU.

TEST INFRASTRUCTURE ONLY.  Nothing is copied into the repository: the reference sources are read where they
lie under /root/reference, rewritten mechanically in memory, written to a temporary directory, compiled, and
the temporary sources are deleted.  Only the shared library lands in oracle/_ref/ (git-ignored).

What runs:
  * the seven GLSL compute shaders (src/vulkansift/shaders/*.comp) compiled as C++ on top of oracle/glsl_emu.h
    after purely syntactic rewrites: `layout(...)` qualifiers and interface blocks become C++ declarations,
    unsuffixed float literals get an `f` (GLSL literals are fp32), `T[N] name` array declarators are reordered,
    signed `%` becomes OpSMod, `main` is renamed;
  * three host functions cut out of the C sources by brace matching: setupGaussianKernels
    (sift_detector.c:52-145), updateScaleSpaceInfo and updateBufferInfo (sift_memory.c:15-87), compiled against
    stub structs that carry exactly the fields they touch.
The dispatch glue (which shader runs on which data) is oracle/ref_driver templates below.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VKSIFT_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libvksift_ref.so")
SHADERS = ["GaussianBlur", "GaussianBlurInterpolated", "DifferenceOfGaussian", "ExtractKeypoints", "ComputeOrientation",
           "ComputeDescriptors", "Get2NearestNeighbors"]


def available():
    return os.path.isdir(os.path.join(REF, "src", "vulkansift", "shaders"))


def _rewrite_shader(src):
    s = src
    s = re.sub(r"#version[^\n]*", "", s)
    s = re.sub(r"(#define\s+PI\s+[0-9.]+)\b", r"\1f", s)
    # GLSL floating literals are fp32
    s = re.sub(r"(?<![\w.])(\d+\.\d*|\.\d+)(?![\w.])", r"\1f", s)
    m = re.search(r"layout\(\s*local_size_x\s*=\s*(\w+)(?:\s*,\s*local_size_y\s*=\s*(\w+))?\s*\)\s*in\s*;", s)
    local = (m.group(1), m.group(2) or "1")
    s = s.replace(m.group(0), "")
    s = re.sub(r"layout\([^)]*\)\s*uniform\s+(sampler2DArray|image2DArray)\s+(\w+)\s*;", r"static \1 \2;", s)
    s = re.sub(r"layout\(push_constant\)\s*uniform\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", r"struct \1 {\2}; static \1 \3;", s, flags=re.S)

    macros = []

    def buffer_block(mm):
        name, body, inst = mm.group(1), mm.group(2), mm.group(3)
        members = re.findall(r"(\w+)\s*(?:\[\s*\])?\s*;", body)
        body = re.sub(r"\[\s*\]", "[1]", body)  # run-time sized array -> trailing array of one ("struct hack")
        out = "struct %s_T {%s}; static %s_T *%s_p;\n" % (name, body, name, name)
        if inst:
            out += "#define %s (*%s_p)\n" % (inst, name)
            macros.append(inst)
        else:
            for member in members:
                out += "#define %s (%s_p->%s)\n" % (member, name, member)
                macros.append(member)
        return out

    s = re.sub(r"layout\(std430,\s*binding\s*=\s*\d+\)\s*buffer\s+(\w+)\s*\{(.*?)\}\s*(\w*)\s*;", buffer_block, s, flags=re.S)
    s = re.sub(r"\bshared\s+(\w+)\[(\w+)\]\s+(\w+)\s*;", r"static \1 \3[\2];", s)
    s = re.sub(r"\bshared\b", "static", s)
    # signed % -> OpSMod.  First "(expr) % x" with one level of nested parentheses, then "a % b".
    s = re.sub(r"\(((?:[^()]|\([^()]*\))+)\)\s*%\s*(\w+)", r"glsl_mod((\1), \2)", s)
    s = re.sub(r"\b(\w+)\s*%\s*(\w+)\b", r"glsl_mod(\1, \2)", s)
    s = s.replace("void main()", "void shader_main()")
    return s, local, macros


WRAP = {
    "blur": r'''
extern "C" void ref_%(fn)s(const float *in, int w, int h, float *out, const float *kernel20, unsigned ksize, int vertical)
{
  img_input.data = in; img_input.w = w; img_input.h = h; img_input.layers = 1;
  img_output.data = out; img_output.w = w; img_output.h = h; img_output.layers = 1;
  push_const.verticalPassFlag = (uint)vertical; push_const.array_layer = 0; push_const.kernel_size = ksize;
  for (int i = 0; i < 20; i++) push_const.kernel[i] = kernel20[i];
  dispatch(shader_main, (w + 7) / 8, (h + 7) / 8, 1, %(lx)s, %(ly)s, 1, false);
}
''',
    "DifferenceOfGaussian": r'''
extern "C" void ref_dog(float *gauss, int w, int h, int n_gauss_layers, float *dog)
{
  img_input.data = gauss; img_input.w = w; img_input.h = h; img_input.layers = n_gauss_layers;
  img_output.data = dog; img_output.w = w; img_output.h = h; img_output.layers = n_gauss_layers - 1;
  dispatch(shader_main, (w + 7) / 8, (h + 7) / 8, n_gauss_layers - 1, %(lx)s, %(ly)s, 1, false);
}
''',
    "ExtractKeypoints": r'''
extern "C" unsigned ref_extract(float *dog, int w, int h, int ns, int octave_idx, float seed_sigma, float dog_thr, float edge_thr,
                                unsigned max_feat, void *feats_out)
{
  dog_input.data = dog; dog_input.w = w; dog_input.h = h; dog_input.layers = ns + 2;
  std::vector<char> mem(8 + (size_t)max_feat * sizeof(SIFT_Feat));
  SIFT_buffer_p = (SIFT_buffer_T *)mem.data();
  IndispatchBuffer_T ind = {0, 1, 1};
  IndispatchBuffer_p = &ind;
  SIFT_buffer_p->nb_elem = 0; SIFT_buffer_p->max_nb_feat = max_feat;
  push_const.octave_idx = octave_idx; push_const.seed_scale_sigma = seed_sigma; push_const.dog_threshold = dog_thr;
  push_const.edge_threshold = edge_thr;
  dispatch(shader_main, (w + 7) / 8, (h + 7) / 8, ns, %(lx)s, %(ly)s, 1, false);
  const unsigned found = SIFT_buffer_p->nb_elem, kept = found < max_feat ? found : max_feat;
  memcpy(feats_out, mem.data() + 8, (size_t)kept * sizeof(SIFT_Feat));
  return found;
}
''',
    "ComputeOrientation": r'''
extern "C" unsigned ref_orientation(float *gauss, int w, int h, int n_layers, void *feats_inout, unsigned n, unsigned max_feat, unsigned max_ori)
{
  octave_input.data = gauss; octave_input.w = w; octave_input.h = h; octave_input.layers = n_layers;
  std::vector<char> mem(8 + (size_t)max_feat * sizeof(SIFT_Feat));
  SIFT_buffer_p = (SIFT_buffer_T *)mem.data();
  IndispatchBuffer_T ind = {n, 1, 1};
  IndispatchBuffer_p = &ind;
  SIFT_buffer_p->nb_elem = n; SIFT_buffer_p->max_nb_feat = max_feat;
  memcpy(mem.data() + 8, feats_inout, (size_t)n * sizeof(SIFT_Feat));
  push_const.max_nb_orientation = max_ori;
  dispatch(shader_main, n, 1, 1, %(lx)s, 1, 1, true);
  const unsigned found = SIFT_buffer_p->nb_elem, kept = found < max_feat ? found : max_feat;
  memcpy(feats_inout, mem.data() + 8, (size_t)kept * sizeof(SIFT_Feat));
  return found;
}
''',
    "ComputeDescriptors": r'''
extern "C" void ref_descriptor(float *gauss, int w, int h, int n_layers, void *feats_inout, unsigned n, unsigned use_vlfeat)
{
  octave_input.data = gauss; octave_input.w = w; octave_input.h = h; octave_input.layers = n_layers;
  std::vector<char> mem(8 + (size_t)(n + 1) * sizeof(SIFT_Feat));
  SIFT_buffer_p = (SIFT_buffer_T *)mem.data();
  SIFT_buffer_p->nb_elem = n; SIFT_buffer_p->max_nb_feat = n;
  memcpy(mem.data() + 8, feats_inout, (size_t)n * sizeof(SIFT_Feat));
  push_const.use_vlfeat_format = use_vlfeat;
  dispatch(shader_main, n, 1, 1, %(lx)s, 1, 1, true);
  memcpy(feats_inout, mem.data() + 8, (size_t)n * sizeof(SIFT_Feat));
}
''',
    "Get2NearestNeighbors": r'''
extern "C" void ref_match(const void *feats_a, unsigned na, const void *feats_b, unsigned nb, void *matches_out)
{
  std::vector<char> ma(8 + (size_t)(na + 1) * sizeof(SIFT_Feat)), mb(8 + (size_t)(nb + 2) * sizeof(SIFT_Feat));
  std::vector<char> md((size_t)(na + 1) * sizeof(SIFT_2NN_Info));
  SIFT_buffer_A_p = (SIFT_buffer_A_T *)ma.data(); SIFT_buffer_B_p = (SIFT_buffer_B_T *)mb.data(); dist_buffer_p = (dist_buffer_T *)md.data();
  SIFT_buffer_A_p->nb_elem = na; SIFT_buffer_A_p->max_nb_elem = na; SIFT_buffer_B_p->nb_elem = nb; SIFT_buffer_B_p->max_nb_elem = nb;
  memcpy(ma.data() + 8, feats_a, (size_t)na * sizeof(SIFT_Feat));
  memcpy(mb.data() + 8, feats_b, (size_t)nb * sizeof(SIFT_Feat));
  dispatch(shader_main, (na + 63) / 64, 1, 1, %(lx)s, 1, 1, false);
  memcpy(matches_out, md.data(), (size_t)na * sizeof(SIFT_2NN_Info));
}
''',
}


def _cut_function(src, signature):
    i = src.index(signature)
    j = src.index("{", i)
    depth, k = 0, j
    while True:
        c = src[k]
        depth += (c == "{") - (c == "}")
        k += 1
        if depth == 0:
            break
    return src[i:k]


HOST_STUBS = r'''
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "vulkansift/vulkansift_types.h"
#define VKSIFT_DETECTOR_MAX_GAUSSIAN_KERNEL_SIZE 20u
#define logDebug(...) ((void)0)
#define logWarning(...) ((void)0)
#define logError(...) ((void)0)
static const char LOG_TAG[] = "ref";
typedef uint64_t VkDeviceSize;
typedef struct { uint32_t width, height; } vksift_OctaveResolution;
typedef struct { bool is_packed; uint32_t curr_input_width, curr_input_height; uint32_t *octave_section_max_nb_feat_arr;
                 VkDeviceSize *octave_section_offset_arr; VkDeviceSize *octave_section_size_arr; } vksift_SiftBufferInfo;
typedef struct { struct { struct { VkDeviceSize minStorageBufferOffsetAlignment; } limits; } physical_device_props; } *vkenv_Device;
typedef struct vksift_SiftMemory_T { vkenv_Device device; vksift_SiftBufferInfo *sift_buffers_info; uint32_t curr_input_image_width,
  curr_input_image_height, curr_nb_octaves; vksift_OctaveResolution *octave_resolutions; uint32_t max_nb_octaves, nb_scales_per_octave,
  max_nb_sift_per_buffer; bool use_upsampling; } *vksift_SiftMemory;
typedef struct vksift_SiftDetector_T { vksift_SiftMemory mem; float *gaussian_kernels; uint32_t *gaussian_kernel_sizes; float input_blur_level,
  seed_scale_sigma; bool use_hardware_interp_kernel; } *vksift_SiftDetector;
'''

HOST_WRAP = r'''
void ref_host_kernel_table(uint32_t ns, int upsample, float input_blur, float seed_sigma, int interp, uint32_t *ksize, float *k)
{
  struct vksift_SiftMemory_T mem; memset(&mem, 0, sizeof(mem));
  mem.nb_scales_per_octave = ns; mem.use_upsampling = upsample != 0;
  struct vksift_SiftDetector_T det; memset(&det, 0, sizeof(det));
  det.mem = &mem; det.input_blur_level = input_blur; det.seed_scale_sigma = seed_sigma; det.use_hardware_interp_kernel = interp != 0;
  setupGaussianKernels(&det);
  memcpy(ksize, det.gaussian_kernel_sizes, sizeof(uint32_t) * (ns + 3));
  memcpy(k, det.gaussian_kernels, sizeof(float) * 20 * (ns + 3));
  free(det.gaussian_kernels); free(det.gaussian_kernel_sizes);
}
uint32_t ref_host_octaves(uint32_t w, uint32_t h, int upsample, uint32_t max_octaves, uint32_t *ow, uint32_t *oh)
{
  struct vksift_SiftMemory_T mem; memset(&mem, 0, sizeof(mem));
  vksift_OctaveResolution res[64];
  mem.octave_resolutions = res; mem.curr_input_image_width = w; mem.curr_input_image_height = h; mem.max_nb_octaves = max_octaves;
  mem.use_upsampling = upsample != 0;
  updateScaleSpaceInfo(&mem);
  for (uint32_t i = 0; i < mem.curr_nb_octaves; i++) { ow[i] = res[i].width; oh[i] = res[i].height; }
  return mem.curr_nb_octaves;
}
void ref_host_sections(uint32_t max_feats, uint32_t n_oct, uint32_t *caps)
{
  struct vksift_SiftMemory_T mem; memset(&mem, 0, sizeof(mem));
  struct { struct { struct { VkDeviceSize minStorageBufferOffsetAlignment; } limits; } physical_device_props; } dev;
  dev.physical_device_props.limits.minStorageBufferOffsetAlignment = 256;
  mem.device = (vkenv_Device)&dev;
  vksift_SiftBufferInfo info; memset(&info, 0, sizeof(info));
  uint32_t c[64]; VkDeviceSize off[64], sz[64];
  info.octave_section_max_nb_feat_arr = c; info.octave_section_offset_arr = off; info.octave_section_size_arr = sz;
  mem.sift_buffers_info = &info; mem.max_nb_octaves = n_oct; mem.curr_nb_octaves = n_oct; mem.max_nb_sift_per_buffer = max_feats;
  updateBufferInfo(&mem, 0);
  memcpy(caps, c, sizeof(uint32_t) * n_oct);
}
'''


def build(force=False, verbose=False):
    if not available():
        raise RuntimeError("reference sources not found under %s" % REF)
    srcs = [os.path.join(REF, "src", "vulkansift", "shaders", n + ".comp") for n in SHADERS]
    srcs += [os.path.join(REF, "src", "vulkansift", f) for f in ("sift_detector.c", "sift_memory.c")]
    deps = srcs + [os.path.join(HERE, f) for f in ("glsl_emu.h", "ref_runtime.cpp", "build_ref.py")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="vksift_ref_")
    try:
        objs = []
        inc = ["-I" + HERE, "-I" + os.path.join(HERE, "..", "include")]
        cxx = ["g++", "-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-w", "-fpermissive"] + inc
        for name in SHADERS:
            text, (lx, ly), macros = _rewrite_shader(open(os.path.join(REF, "src", "vulkansift", "shaders", name + ".comp")).read())
            key = "blur" if name.startswith("GaussianBlur") else name
            fn = {"GaussianBlur": "blur_plain", "GaussianBlurInterpolated": "blur_interp"}.get(name, name)
            undef = "".join("#undef %s\n" % m for m in macros if key in ("ExtractKeypoints", "ComputeOrientation", "ComputeDescriptors"))
            # the wrappers of the unnamed-block shaders address the buffers through the *_p pointers only
            cpp = ('#include "glsl_emu.h"\nnamespace ref_%s {\nusing namespace glsl;\n%s\n%s%s\n}\n' %
                   (name, text, undef, WRAP[key] % {"fn": fn, "lx": lx, "ly": ly}))
            path = os.path.join(tmp, name + ".cpp")
            open(path, "w").write(cpp)
            obj = os.path.join(tmp, name + ".o")
            subprocess.run(cxx + ["-c", path, "-o", obj], check=True, capture_output=not verbose)
            objs.append(obj)
        det = open(os.path.join(REF, "src", "vulkansift", "sift_detector.c")).read()
        mem = open(os.path.join(REF, "src", "vulkansift", "sift_memory.c")).read()
        host = HOST_STUBS + _cut_function(det, "static void setupGaussianKernels(vksift_SiftDetector detector)") + "\n" + \
            _cut_function(mem, "void updateScaleSpaceInfo(vksift_SiftMemory memory)") + "\n" + \
            _cut_function(mem, "void updateBufferInfo(vksift_SiftMemory memory, uint32_t buffer_idx)") + "\n" + HOST_WRAP
        path = os.path.join(tmp, "host.c")
        open(path, "w").write(host)
        obj = os.path.join(tmp, "host.o")
        subprocess.run(["gcc", "-O1", "-std=gnu11", "-fPIC", "-ffp-contract=off", "-w"] + inc + ["-c", path, "-o", obj], check=True,
                       capture_output=not verbose)
        objs.append(obj)
        obj = os.path.join(tmp, "runtime.o")
        subprocess.run(cxx + ["-c", os.path.join(HERE, "ref_runtime.cpp"), "-o", obj], check=True, capture_output=not verbose)
        objs.append(obj)
        subprocess.run(["g++", "-shared", "-o", OUT] + objs + ["-lm"], check=True, capture_output=not verbose)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)  # no reference-derived source is kept
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
