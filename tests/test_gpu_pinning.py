"""gpu-marked aliases of the oracle-pinning tests.

The driver's round-end run on the GPU box selects `-m gpu` only, which deselects the CPU tests that pin the oracle
against the reference's own code (tests/test_ref_parity.py, tests/test_ref_chain.py, tests/test_oracle.py::golden
checks).  They need nothing but the committed tests/golden/ref_vectors.npz (and, for the live / chained ones, the
prebuilt oracle/_ref library that travels with the snapshot), so the same functions run here under the gpu marker:
the record of the GPU run then shows the checker was checked on the same box, in the same session, as the parity tests
that rely on it.
"""
import pytest

import test_ref_chain as _chain
import test_ref_parity as _parity
from test_ref_parity import gold  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def test_pinning_host_tables_equal_reference_functions(gold, oracle_mod):  # noqa: F811
    _parity.test_host_tables_equal_reference_functions(gold, oracle_mod)


def test_pinning_shader_stages_equal_reference_shaders(gold, oracle_mod):  # noqa: F811
    _parity.test_shader_stages_equal_reference_shaders(gold, oracle_mod)


def test_pinning_matcher_equals_reference_shader(gold, oracle_mod):  # noqa: F811
    _parity.test_matcher_equals_reference_shader(gold, oracle_mod)


def test_pinning_live_reference_harness(oracle_mod):
    _parity.test_live_reference_harness_more_configurations(oracle_mod)


def test_pinning_chained_reference_shaders_on_c1(oracle_mod, c1_image):
    _chain.test_chained_reference_shaders_on_c1(oracle_mod, c1_image)
