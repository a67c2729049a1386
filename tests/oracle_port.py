"""ctypes driver for the plain-C restatement oracle/bfv_oracle.c (oracle/_ref/libbfv_oracle.so) — test infrastructure."""
import ctypes as C
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_SO = os.path.join(_ROOT, "oracle", "_ref", "libbfv_oracle.so")
u64 = C.c_uint64
vp = C.c_void_p
ORC_MAXK = 64


class OrcCtx(C.Structure):
    _fields_ = [("n", C.c_size_t), ("logn", C.c_int), ("K", C.c_int), ("q", u64 * ORC_MAXK), ("t", u64), ("m_sk", u64),
                ("gamma", u64), ("m_tilde", u64), ("aux", u64 * (ORC_MAXK + 4))]


def _p(a):
    return a.ctypes.data_as(vp)


class OraclePort:
    def __init__(self, path=PORT_SO):
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL)
        L = self.lib
        L.orc_fnv1a64.restype = u64
        L.orc_fnv1a64.argtypes = [vp, C.c_size_t]
        L.orc_min_root.restype = u64
        L.orc_min_root.argtypes = [u64, C.c_size_t]
        L.orc_is_prime.argtypes = [u64]
        L.orc_ntt_forward.argtypes = [vp, C.c_size_t, u64]
        L.orc_ntt_inverse.argtypes = [vp, C.c_size_t, u64]
        L.orc_ntt_forward.restype = None
        L.orc_ntt_inverse.restype = None
        L.orc_galois_elt_from_step.restype = C.c_uint32

    def fnv(self, words):
        w = np.ascontiguousarray(words, dtype=np.uint64).reshape(-1)
        return int(self.lib.orc_fnv1a64(_p(w), w.size))

    def min_root(self, p, n):
        return int(self.lib.orc_min_root(u64(p), n))

    def ntt_forward(self, poly, p):
        a = np.ascontiguousarray(poly, dtype=np.uint64).copy()
        self.lib.orc_ntt_forward(_p(a), a.size, u64(p))
        return a

    def ntt_inverse(self, poly, p):
        a = np.ascontiguousarray(poly, dtype=np.uint64).copy()
        self.lib.orc_ntt_inverse(_p(a), a.size, u64(p))
        return a

    def context(self, n, moduli, t):
        return PortContext(self, n, moduli, t)


class PortContext:
    def __init__(self, port, n, moduli, t):
        self.P = port
        self.L = port.lib
        self.n = n
        self.moduli = [int(m) for m in moduli]
        self.K = len(moduli)
        self.k = self.K - 1 if self.K > 1 else 1
        self.c = OrcCtx()
        arr = (u64 * self.K)(*self.moduli)
        rc = self.L.orc_ctx_init(C.byref(self.c), C.c_size_t(n), arr, self.K, u64(t))
        assert rc == 0, rc

    def base_b_size(self, k=None):
        return int(self.L.orc_base_b_size(C.byref(self.c), k or self.k))

    def aux_primes(self):
        return [int(x) for x in self.c.aux[: self.K + 3]]

    def _ct(self, arr):
        return np.ascontiguousarray(arr, dtype=np.uint64)

    def add(self, a, b):
        a, b = self._ct(a), self._ct(b)
        out = np.empty_like(a)
        self.L.orc_add(C.byref(self.c), self.k, _p(a), _p(b), _p(out), a.shape[0])
        return out

    def sub(self, a, b):
        a, b = self._ct(a), self._ct(b)
        out = np.empty_like(a)
        self.L.orc_sub(C.byref(self.c), self.k, _p(a), _p(b), _p(out), a.shape[0])
        return out

    def negate(self, a):
        a = self._ct(a)
        out = np.empty_like(a)
        self.L.orc_negate(C.byref(self.c), self.k, _p(a), _p(out), a.shape[0])
        return out

    def multiply(self, a, b):
        a, b = self._ct(a), self._ct(b)
        out = np.zeros((a.shape[0] + b.shape[0] - 1, self.k, self.n), dtype=np.uint64)
        rc = self.L.orc_multiply(C.byref(self.c), self.k, _p(a), a.shape[0], _p(b), b.shape[0], _p(out))
        assert rc == 0
        return out

    def relinearize(self, in3, key):
        in3, key = self._ct(in3), self._ct(key)
        out = np.zeros((2, self.k, self.n), dtype=np.uint64)
        assert self.L.orc_relinearize(C.byref(self.c), self.k, _p(in3), _p(key), _p(out)) == 0
        return out

    def apply_galois(self, in2, elt, key):
        in2, key = self._ct(in2), self._ct(key)
        out = np.zeros((2, self.k, self.n), dtype=np.uint64)
        assert self.L.orc_apply_galois(C.byref(self.c), self.k, _p(in2), C.c_uint32(elt), _p(key), _p(out)) == 0
        return out

    def galois_elt_from_step(self, steps):
        return int(self.L.orc_galois_elt_from_step(C.byref(self.c), C.c_int(steps)))

    def multiply_plain(self, a, plain):
        a, plain = self._ct(a), self._ct(plain)
        out = np.zeros_like(a)
        rc = self.L.orc_multiply_plain(C.byref(self.c), self.k, _p(a), a.shape[0], _p(plain), C.c_size_t(plain.size), _p(out))
        assert rc == 0
        return out

    def add_plain(self, a, plain, subtract=False):
        a, plain = self._ct(a), self._ct(plain)
        out = np.zeros_like(a)
        rc = self.L.orc_add_plain(C.byref(self.c), self.k, _p(a), a.shape[0], _p(plain), C.c_size_t(plain.size), _p(out),
                                  int(subtract))
        assert rc == 0
        return out

    def mod_switch_to_next(self, a):
        a = self._ct(a)
        out = np.zeros((a.shape[0], self.k - 1, self.n), dtype=np.uint64)
        assert self.L.orc_mod_switch_to_next(C.byref(self.c), self.k, _p(a), a.shape[0], _p(out)) == 0
        return out

    def decrypt(self, ct, sk_ntt):
        ct, sk = self._ct(ct), self._ct(sk_ntt)
        out = np.zeros(self.n, dtype=np.uint64)
        assert self.L.orc_decrypt(C.byref(self.c), self.k, _p(ct), ct.shape[0], _p(sk), _p(out)) == 0
        return out


class RnsSteps:
    """RNSTool steps with explicit tiny bases, for the reference's own KATs (tests/seal/util/rns.cpp:460-853)."""

    def __init__(self, port, q, n):
        self.L = port.lib
        self.q = [int(x) for x in q]
        self.k = len(q)
        self.n = n
        self.nB = self.k  # t = 0 in the reference tests -> never extended
        buf = (u64 * (self.nB + 1))()
        self.L.orc_aux_bases.restype = u64
        self.gamma = int(self.L.orc_aux_bases(C.c_size_t(n), self.nB, buf))
        self.bsk = [int(x) for x in buf]
        self.m_tilde = 1 << 32
        self._q = (u64 * self.k)(*self.q)
        self._bsk = (u64 * len(self.bsk))(*self.bsk)

    def _run(self, name, arr, rows_out):
        a = np.ascontiguousarray(arr, dtype=np.uint64)
        out = np.zeros(rows_out * self.n, dtype=np.uint64)
        getattr(self.L, name).restype = None
        getattr(self.L, name)(self._q, self.k, self._bsk, len(self.bsk), C.c_size_t(self.n), _p(a), _p(out))
        return [int(x) for x in out]

    def fastbconv_m_tilde(self, arr):
        return self._run("orc_fastbconv_m_tilde", arr, len(self.bsk) + 1)

    def sm_mrq(self, arr):
        return self._run("orc_sm_mrq", arr, len(self.bsk))

    def fast_floor(self, arr):
        return self._run("orc_fast_floor", arr, len(self.bsk))

    def fastbconv_sk(self, arr):
        return self._run("orc_fastbconv_sk", arr, self.k)
