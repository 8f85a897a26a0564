"""Run what ships: the C / C++ callers under examples/ are compiled against include/ + libvulkansift.so exactly as
INTEGRATION.md tells a maintainer to, and EXECUTED on the GPU (tests/test_abi.py only link-tests them on the CPU box)."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(ROOT, "vulkansift_b200", "lib")


def _build(tmp_path, name, cxx=False):
    src = os.path.join(ROOT, "examples", name + (".cpp" if cxx else ".c"))
    exe = str(tmp_path / name)
    cmd = (["g++", "-std=c++17", "-O1"] if cxx else ["gcc", "-std=c11", "-O2"]) + [src, "-I" + os.path.join(ROOT, "include"), "-L" + LIB_DIR,
                                                                                "-lvulkansift", "-Wl,-rpath," + LIB_DIR, "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_plain_c_detect_match_example_runs(tmp_path):
    r = subprocess.run([_build(tmp_path, "detect_match")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"features: (\d+) / (\d+), matches passing the ratio test: (\d+) of (\d+)", r.stdout)
    assert m, r.stdout
    na, nb, good, nm = map(int, m.groups())
    assert na > 20 and nb > 20 and nm == na and good > na // 4  # the second image is the first one shifted by 6 px


def test_perf_runtime_driver_runs(tmp_path):
    """The port of the reference's perf_sift_runtime protocol (src/perf/perf_runtime.cpp): 50 + 500 iterations, result file."""
    r = subprocess.run([_build(tmp_path, "perf_runtime")], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    out = open(tmp_path / "runtime_results_vulkansift.txt").read().strip()
    mean_ms, n_feat = out.split(";")
    assert 0.05 < float(mean_ms) < 50.0 and int(n_feat) > 1000, out


def test_perf_matching_driver_runs(tmp_path):
    """The port of the reference's perf_sift_match protocol (src/perf/perf_matching.cpp, perf_common.cpp:109-170)."""
    r = subprocess.run([_build(tmp_path, "perf_matching")], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"matches: (\d+) filtered of (\d+); inliers within 2.5 px of the known homography: (\d+)", r.stdout)
    assert m, r.stdout
    filt, total, inl = map(int, m.groups())
    assert total > 100 and filt > total // 4 and inl > 0.8 * filt


def test_cpp_caller_throws_through_the_error_callback(tmp_path):
    """src/examples/test_sift_error_handling.cpp:6-16: the callback throws, the exception crosses the C entry point."""
    r = subprocess.run([_build(tmp_path, "error_handling", cxx=True)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "buffer 5: std::invalid_argument caught" in r.stdout and "error_handling: ok" in r.stdout, r.stdout


def test_exchange_example_three_processes_from_plain_c(tmp_path):
    """examples/exchange_pairs.c: three forked workers (pipes for the IPC handles and barriers, no MPI, no torch) detect on their own
    image, exchange descriptors through vksiftx_exchangeMatchAllPeers and check the records they get for every peer."""
    r = subprocess.run([_build(tmp_path, "exchange_pairs")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "exchange_pairs: ok" in r.stdout, r.stdout
    assert len(re.findall(r"rank \d \(\d+ features\) vs rank \d \(\d+ features\): \d+ matches passing the ratio test", r.stdout)) == 6, r.stdout
