"""The seal_fhe crate's unit tests (restated in tests/seal_fhe_crate_tests.py) on the CPU emulation build."""
import seal_fhe_crate_tests as crate
from sunscreen_b200 import seal_fhe


def test_bfv_evaluator_crate_tests(emu_lib):
    seal_fhe.use_library(emu_lib.lib)
    done = crate.all_tests()
    assert len(done) == 12


def test_lane_overflow_assumption(emu_lib):
    seal_fhe.use_library(emu_lib.lib)
    crate.lane_overflow_assumption()


def test_serialization_components_polyarray_crate_tests(emu_lib):
    seal_fhe.use_library(emu_lib.lib)
    assert len(crate.serialization_and_components_tests()) == 4
