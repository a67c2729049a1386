"""The seal_fhe crate's unit tests (tests/seal_fhe_crate_tests.py) on the real CUDA library."""
import pytest

import seal_fhe_crate_tests as crate
from sunscreen_b200 import seal_fhe

pytestmark = pytest.mark.gpu


def test_bfv_evaluator_crate_tests():
    from sunscreen_b200.lib import B200Lib
    seal_fhe.use_library(B200Lib.default().lib)
    assert len(crate.all_tests()) == 12


def test_lane_overflow_assumption():
    from sunscreen_b200.lib import B200Lib
    seal_fhe.use_library(B200Lib.default().lib)
    crate.lane_overflow_assumption()


def test_serialization_components_polyarray_crate_tests():
    from sunscreen_b200.lib import B200Lib
    seal_fhe.use_library(B200Lib.default().lib)
    assert len(crate.serialization_and_components_tests()) == 4
