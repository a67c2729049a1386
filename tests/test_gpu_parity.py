"""GPU parity (pytest -m gpu): the CUDA library on cuda:0, called through the C ABI, against the unmodified
reference (prebuilt oracle/_ref/libsealc_ref.so) — every output uint64 must be equal."""
import numpy as np
import pytest

import parity_checks as pc
from params import PARAMS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    from backends import CudaBackend
    return CudaBackend()


@pytest.fixture(scope="module", params=["n4096", "n8192", "n8192_54", "n8192_49", "n16384", "n32768"])
def pair(request, be, ref):
    n, moduli, t = PARAMS[request.param]
    return pc.Pair(be, n, moduli, t)


def test_native_library_loaded(be):
    import os
    assert os.path.basename(be.lib.path) == "libb200bfv.so"
    assert be.lib.device_count() >= 1


def test_context_constants(pair):
    pc.check_context(pair)


def test_ntt(pair):
    pc.check_ntt(pair, items=5)


def test_elementwise(pair):
    pc.check_elementwise(pair)


def test_multiply_and_relinearize(pair):
    m3, rm = pc.check_multiply(pair, with_sizes=pair.n <= 16384)
    pc.check_relin(pair, m3, rm)


def test_galois(pair):
    if not pair.ctx.using_batching:
        pytest.skip("t does not support batching (reference raises logic_error)")
    pc.check_galois(pair)


def test_plain_ops(pair):
    pc.check_plain(pair)


def test_mod_switch(pair):
    pc.check_modswitch(pair)


def test_host_buffer_pipeline(pair):
    if pair.n > 16384:
        pytest.skip("covered at n <= 16384")
    pc.check_host_pipeline(pair, batch=7)


def test_batch_strides(pair):
    pc.check_batch(pair, batch=5 if pair.n <= 16384 else 2)


def test_small_kernels(pair):
    pc.check_small_kernels(pair)


def test_noise_norm(pair):
    pc.check_noise_norm(pair)


def test_encrypted_roundtrip(pair):
    if not pair.ctx.using_batching:
        pytest.skip("needs a batching plain modulus")
    pc.check_encrypted_roundtrip(pair)


def test_full_size_properties(be):
    """BASELINE config 2/3 sizes through size-independent properties: NTT round trip over 4096x4 polynomials,
    and multiply_relin over a 256-item batch equal to the same items computed one by one."""
    n, moduli, t = PARAMS["n8192"]
    from sunscreen_b200.lib import B200Context
    ctx = B200Context(n, moduli, t)
    k = ctx.k()
    rng = np.random.default_rng(11)
    items = 4096
    x = pc.rand_ct(rng, moduli, k, n, size=1, batch=items)[:, 0]
    d = be.to_dev(x)
    ctx.ntt_forward(d, items)
    f = be.to_host(d)
    assert not np.array_equal(f, x)
    ctx.ntt_inverse(d, items)
    pc.eq(be.to_host(d), x, "4096x4 NTT round trip")
    # linearity of the forward transform: NTT(a+b) = NTT(a)+NTT(b) mod q
    y = pc.rand_ct(rng, moduli, k, n, size=1, batch=8)[:, 0]
    s = np.stack([(x[:8, i].astype(object) + y[:, i].astype(object)) % moduli[i] for i in range(k)], axis=1).astype(np.uint64)
    ds, dy = be.to_dev(s), be.to_dev(y)
    ctx.ntt_forward(ds, 8)
    ctx.ntt_forward(dy, 8)
    fy = be.to_host(dy)
    fs = np.stack([(f[:8, i].astype(object) + fy[:, i].astype(object)) % moduli[i] for i in range(k)], axis=1).astype(np.uint64)
    pc.eq(be.to_host(ds), fs, "NTT linearity")
    B = 256
    A = pc.rand_ct(rng, moduli, k, n, batch=B)
    Bc = pc.rand_ct(rng, moduli, k, n, batch=B)
    key = pc.rand_ksk(rng, moduli, k, n)
    dA, dB, dK = be.to_dev(A), be.to_dev(Bc), be.to_dev(key)
    o = be.empty((B, 2, k, n))
    ctx.multiply_relin(dA, dB, dK, o, B)
    got = be.to_host(o)
    for i in (0, 1, 100, 255):
        oi = be.empty((2, k, n))
        ctx.multiply_relin(be.to_dev(A[i]), be.to_dev(Bc[i]), dK, oi, 1)
        pc.eq(got[i], be.to_host(oi), f"batched item {i} == single")
    # host-buffer entry point returns the same words
    oh = np.zeros((B, 2, k, n), dtype=np.uint64)
    ctx.multiply_relin_host(A, Bc, dK, oh, B)
    pc.eq(oh, got, "multiply_relin_host")
