"""CPU-side parity: the product sources compiled in the test-only emulation mode (tests/emu) against the
unmodified reference, on the RNG-free vectors of SURVEY.md App. B.  Covers host precompute, constant folding,
strides and the orchestration of every layer-1 entry point; the CUDA build itself is covered by test_gpu_parity."""
import pytest

import parity_checks as pc
from backends import EmuBackend
from params import PARAMS


@pytest.fixture(scope="module", params=["n4096", "n8192", "n8192_54", "n8192_49"])
def pair(request, emu_lib, ref):
    n, moduli, t = PARAMS[request.param]
    return pc.Pair(EmuBackend(emu_lib), n, moduli, t)


def test_context_constants(pair):
    pc.check_context(pair)


def test_ntt(pair):
    pc.check_ntt(pair)


def test_elementwise(pair):
    pc.check_elementwise(pair)


def test_multiply_and_relinearize(pair):
    m3, rm = pc.check_multiply(pair)
    pc.check_relin(pair, m3, rm)


def test_galois(pair):
    if not pair.ctx.using_batching:
        pytest.skip("t does not support batching (reference raises logic_error)")
    pc.check_galois(pair)


def test_plain_ops(pair):
    pc.check_plain(pair)


def test_mod_switch(pair):
    pc.check_modswitch(pair)


def test_batch_strides(pair):
    if pair.n > 4096:
        pytest.skip("batch stride test runs on the smallest set only (emulation speed)")
    pc.check_batch(pair)


def test_host_buffer_pipeline(pair):
    if pair.n > 4096:
        pytest.skip("host pipeline test runs on the smallest set only (emulation speed)")
    pc.check_host_pipeline(pair)


def test_encrypted_roundtrip(pair):
    if not pair.ctx.using_batching:
        pytest.skip("needs a batching plain modulus")
    pc.check_encrypted_roundtrip(pair)


def test_n32768_two_level_transform(emu_lib, ref):
    """BASELINE config 5 parameters (n=32768, 15 data residues + special): the two-level NTT and K=15 kernels."""
    n, moduli, t = PARAMS["n32768"]
    P = pc.Pair(EmuBackend(emu_lib), n, moduli, t)
    pc.check_context(P)
    pc.check_ntt(P, items=1)
    m3, rm = pc.check_multiply(P, with_sizes=False)
    pc.check_relin(P, m3, rm)
    pc.check_galois(P)
    pc.check_plain(P)
    pc.check_modswitch(P)
