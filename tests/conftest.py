import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip silently; but a plain run skips gpu tests.
    markexpr = config.getoption("-m") or ""
    if "gpu" in markexpr and "not gpu" not in markexpr:
        return
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def c1_image():
    from vulkansift_b200.synth import blob_image, C1
    return blob_image(**C1)


def assert_features_equal(got, exp, what=""):
    """Bit-exact comparison of two vksift_Feature arrays with a readable diff."""
    assert len(got) == len(exp), "%s feature count %d != oracle %d" % (what, len(got), len(exp))
    for name in got.dtype.names:
        a, b = got[name], exp[name]
        if a.dtype.kind == "f":
            same = a.view(np.uint32) == b.view(np.uint32)
        else:
            same = a == b
        if not np.all(same):
            bad = np.argwhere(~same)
            i = int(bad[0][0])
            raise AssertionError("%s field %s differs at %d of %d rows; first at %s: got %r expected %r" %
                                 (what, name, len(np.unique(bad[:, 0])), len(got), bad[0].tolist(), got[name][i], exp[name][i]))
