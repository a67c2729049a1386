"""Checks of the SEAL-named ABI layer shared by the CPU (emu) and GPU test files."""
import ctypes as C

import numpy as np

import refseal
from parity_checks import eq
from sealc_driver import Sealc, SealcError

vp, u64 = C.c_void_p, C.c_uint64


def simple_multiply_sequence(S, n, moduli, t, seed=1):
    """examples/simple_multiply through the FFI: keys + fresh encryptions from the reference,
    Evaluator_Multiply (size 3) -> Evaluator_Relinearize (size 2) -> Decryptor_InvariantNoiseBudget -> Decryptor_Decrypt."""
    R = refseal.RefContext(n, moduli, t)
    O = S.context(n, moduli, t)
    assert O.parameters_set
    assert list(O.first_id) == list(R.first_parms_id) and list(O.key_id) == list(R.key_parms_id)
    kg = R.keygen()
    sk, pk, rk = R.secret_key(kg), R.public_key(kg), R.relin_keys(kg)
    enc, dec = R.encryptor(pk), R.decryptor(sk)
    ca, cb = R.encrypt(enc, R.new_pt(np.array([15], dtype=np.uint64))), R.encrypt(enc, R.new_pt(np.array([5], dtype=np.uint64)))
    oa, ob = O.new_ct(R.ct_words(ca)), O.new_ct(R.ct_words(cb))
    ork = O.new_ksk(R.ksk_words(rk))
    rm, om = R.multiply(ca, cb), O.multiply(oa, ob)
    eq(O.ct_words(om), R.ct_words(rm), "Evaluator_Multiply words")
    rr, orr = R.relinearize(rm, rk), O.relinearize(om, ork)
    eq(O.ct_words(orr), R.ct_words(rr), "Evaluator_Relinearize words")
    # decrypt through OUR Decryptor with the reference's secret key words
    h = vp()
    R.ref.call("SecretKey_Data", sk, C.byref(h))
    skw = R.pt_coeffs(h)
    odec = O.decryptor(skw)
    assert O.noise_budget(odec, orr) == R.noise_budget(dec, rr) > 0
    got = O.pt_coeffs(O.decrypt(odec, orr))
    exp = R.pt_coeffs(R.decrypt(dec, rr))
    eq(got, exp, "Decryptor_Decrypt")
    assert int(got[0]) == 75 and got.size == 1


def evaluator_surface(S, n, moduli, t):
    """Every Evaluator entry point seal_fhe uses, on the RNG-free App. B vectors, against the reference's words."""
    R = refseal.RefContext(n, moduli, t)
    O = S.context(n, moduli, t)
    inp = refseal.appendix_b_inputs(n, moduli, t)
    ra, rb, rp = R.new_ct(inp["a"]), R.new_ct(inp["b"]), R.new_pt(inp["p"])
    oa, ob, op = O.new_ct(inp["a"]), O.new_ct(inp["b"]), O.new_pt(inp["p"])
    rrlk, orlk = R.new_ksk({0: inp["rlk"]}), O.new_ksk({0: inp["rlk"]})
    W = lambda h: O.ct_words(h)
    eq(W(O.add(oa, ob)), R.ct_words(R.add(ra, rb)), "Add")
    eq(W(O.sub(oa, ob)), R.ct_words(R.sub(ra, rb)), "Sub")
    eq(W(O.negate(oa)), R.ct_words(R.negate(ra)), "Negate")
    rm, om = R.multiply(ra, rb), O.multiply(oa, ob)
    eq(W(om), R.ct_words(rm), "Multiply")
    eq(W(O.square(oa)), R.ct_words(R.square(ra)), "Square")
    eq(W(O.relinearize(om, orlk)), R.ct_words(R.relinearize(rm, rrlk)), "Relinearize")
    eq(W(O.add(om, oa)), R.ct_words(R.add(rm, ra)), "Add sizes (3,2)")
    eq(W(O.sub(oa, om)), R.ct_words(R.sub(ra, rm)), "Sub sizes (2,3)")
    eq(W(O.multiply_plain(oa, op)), R.ct_words(R.multiply_plain(ra, rp)), "MultiplyPlain")
    eq(W(O.add_plain(oa, op)), R.ct_words(R.add_plain(ra, rp)), "AddPlain")
    eq(W(O.sub_plain(oa, op)), R.ct_words(R.sub_plain(ra, rp)), "SubPlain")
    eq(W(O.mod_switch_to_next(oa)), R.ct_words(R.mod_switch_to_next(ra)), "ModSwitchToNext1")
    eq(W(O.add_many([oa, ob, oa])), R.ct_words(R.add(R.add(ra, rb), ra)), "AddMany")
    # in-place aliasing (evaluator_base.rs:184-196 passes dest == src)
    oc = O.new_ct(inp["a"])
    O.S.call("Evaluator_Add", O.ev, oc, ob, oc)
    eq(W(oc), R.ct_words(R.add(ra, rb)), "Add in place")
    # multiply_many / exponentiate: same product tree as the reference (S/evaluator.cpp:1535-1643)
    d = R.new_ct()
    arr = (vp * 3)(ra, rb, ra)
    R.ref.call("Evaluator_MultiplyMany", R.ev, u64(3), arr, rrlk, d, None)
    eq(W(O.multiply_many([oa, ob, oa], orlk)), R.ct_words(d), "MultiplyMany")
    d2 = R.new_ct()
    R.ref.call("Evaluator_Exponentiate", R.ev, ra, u64(3), rrlk, d2, None)
    eq(W(O.exponentiate(oa, 3, orlk)), R.ct_words(d2), "Exponentiate")
    if t % (2 * n) == 1:
        gl = {1: inp["glk3"], (2 * n - 2) // 2: inp["glkc"]}
        rg, og = R.new_ksk(gl), O.new_ksk(gl)
        eq(W(O.rotate_rows(oa, 1, og)), R.ct_words(R.rotate_rows(ra, 1, rg)), "RotateRows(1)")
        eq(W(O.rotate_columns(oa, og)), R.ct_words(R.rotate_columns(ra, rg)), "RotateColumns")
        # missing key -> NAF decomposition path: steps=3 = 4 - 1 needs keys for +4 and -1; only +1 present -> error like the reference
        for steps in (3,):
            try:
                R.rotate_rows(ra, steps, rg)
                ref_code = 0
            except refseal.SealError as e:
                ref_code = e.code
            try:
                O.rotate_rows(oa, steps, og)
                our_code = 0
            except SealcError as e:
                our_code = e.code
            assert our_code == ref_code, (hex(our_code), hex(ref_code))


def error_codes(S, R_lib, n, moduli, t):
    """HRESULT parity with the reference on the failure paths Rust maps (seal_fhe/src/error.rs:65-91)."""
    R = refseal.RefContext(n, moduli, t)
    O = S.context(n, moduli, t)
    inp = refseal.appendix_b_inputs(n, moduli, t)
    ra, oa = R.new_ct(inp["a"]), O.new_ct(inp["a"])
    rd, od = R.new_ct(), O._dst()
    E_POINTER, E_INVALIDARG, E_INVOP = 0x80004003, 0x80070057, 0x80131509
    # NULL handle
    assert O.S.rc("Evaluator_Multiply", O.ev, oa, oa, None, None) == R.ref.call_rc("Evaluator_Multiply", R.ev, ra, ra, None, None) == E_POINTER
    assert O.S.rc("Evaluator_Add", None, oa, oa, od) == E_POINTER
    # empty operand
    re, oe = R.new_ct(), O._dst()
    assert O.S.rc("Evaluator_Multiply", O.ev, oe, oa, od, None) == R.ref.call_rc("Evaluator_Multiply", R.ev, re, ra, rd, None) == E_INVALIDARG
    # transparent result: a - a
    assert O.S.rc("Evaluator_Sub", O.ev, oa, oa, od) == R.ref.call_rc("Evaluator_Sub", R.ev, ra, ra, rd) == E_INVOP
    # relinearize with keys of the wrong parms_id / missing keys
    rk_empty, ok_empty = vp(), vp()
    R.ref.call("KSwitchKeys_Create1", C.byref(rk_empty))
    O.S.call("KSwitchKeys_Create1", C.byref(ok_empty))
    rm, om = R.multiply(ra, ra), O.multiply(oa, oa)
    assert O.S.rc("Evaluator_Relinearize", O.ev, om, ok_empty, od, None) == R.ref.call_rc("Evaluator_Relinearize", R.ev, rm, rk_empty, rd, None) == E_INVALIDARG
    # NTT-form input to multiply
    O.S.call("Ciphertext_SetIsNTTForm", oa, C.c_bool(True))
    R.ref.call("Ciphertext_SetIsNTTForm", ra, C.c_bool(True))
    assert O.S.rc("Evaluator_Multiply", O.ev, oa, oa, od, None) == R.ref.call_rc("Evaluator_Multiply", R.ev, ra, ra, rd, None) == E_INVALIDARG
    # invalid parameters: context reports parameters not set, Evaluator_Create fails
    bad = S.context(n, [moduli[0], moduli[0] + 2], t, sec=0)
    assert not bad.parameters_set
    ev = vp()
    assert S.rc("Evaluator_Create", bad.ctx, C.byref(ev)) == E_INVALIDARG


def _ref_pk_words(R, pk):
    h = vp()
    R.ref.call("PublicKey_Data", pk, C.byref(h))
    return R.ct_words_any(h)


def seeded_encryption_parity(S, n, moduli, t):
    """pk-encryption from a fixed 64-byte seed: our B200_Encryptor_EncryptSetSeed reproduces the reference's
    Encryptor_EncryptReturnComponentsSetSeed ciphertext word for word (same Blake2xb stream, same samplers)."""
    R = refseal.RefContext(n, moduli, t)
    O = S.context(n, moduli, t)
    kg = R.keygen()
    pk = R.public_key(kg)
    pkw = _ref_pk_words(R, pk)
    enc_r = R.encryptor(pk)
    # our Encryptor over the same public key words
    opk = vp()
    O.S.call("PublicKey_Create1", C.byref(opk))
    opk_ct = vp()
    O.S.call("PublicKey_Data", opk, C.byref(opk_ct))
    w = np.ascontiguousarray(pkw, dtype=np.uint64)
    O.S.call("B200_Ciphertext_SetWords", opk_ct, O.ctx, O.key_id, u64(2), C.c_bool(True), w.ctypes.data_as(C.POINTER(u64)))
    enc_o = vp()
    O.S.call("Encryptor_Create", O.ctx, opk, None, C.byref(enc_o))
    rng = np.random.default_rng(4)
    for trial, seed in enumerate(([0] * 8, [1, 2, 3, 4, 5, 6, 7, 8], list(rng.integers(0, 2**63, size=8)))):
        msg = rng.integers(0, t, size=n if trial else 3, dtype=np.uint64)
        seed_arr = (u64 * 8)(*[int(x) for x in seed])
        # reference
        rct = R.new_ct()
        pa_u, pa_e, rem = vp(), vp(), vp()
        R.ref.call("PolynomialArray_Create", None, C.byref(pa_u))
        R.ref.call("PolynomialArray_Create", None, C.byref(pa_e))
        R.ref.call("Plaintext_Create1", None, C.byref(rem))
        R.ref.call("Encryptor_EncryptReturnComponentsSetSeed", enc_r, R.new_pt(msg), C.c_bool(False), rct, pa_u, pa_e, rem,
                   seed_arr, None)
        # ours
        oct_ = O._dst()
        O.S.call("B200_Encryptor_EncryptSetSeed", enc_o, O.new_pt(msg), seed_arr, oct_)
        eq(O.ct_words(oct_), R.ct_words(rct), f"seeded encryption, trial {trial}")


def keygen_interop(S, n, moduli, t):
    """Keys made by OUR KeyGenerator (host sampling + GPU arithmetic) are valid keys for the REFERENCE: a ciphertext
    encrypted by our Encryptor is multiplied / relinearized / rotated by the reference's Evaluator with our keys and
    decrypted by the reference's Decryptor with our secret key — and the same through our own layer."""
    R = refseal.RefContext(n, moduli, t)
    O = S.context(n, moduli, t)
    kg = vp()
    O.S.call("KeyGenerator_Create1", O.ctx, C.byref(kg))
    sk, pk, rlk = vp(), vp(), vp()
    O.S.call("KeyGenerator_SecretKey", kg, C.byref(sk))
    O.S.call("KeyGenerator_CreatePublicKey", kg, C.c_bool(False), C.byref(pk))
    O.S.call("KeyGenerator_CreateRelinKeys", kg, C.c_bool(False), C.byref(rlk))
    # secret key words -> reference SecretKey
    skd = vp()
    O.S.call("SecretKey_Data", sk, C.byref(skd))
    skw = O.pt_coeffs(skd)
    rsk = vp()
    R.ref.call("SecretKey_Create1", C.byref(rsk))
    rskd = vp()
    R.ref.call("SecretKey_Data", rsk, C.byref(rskd))
    R.ref.call("Plaintext_Resize", rskd, u64(skw.size))
    C.memmove(R.ref.lib.refshim_pt_data(rskd), skw.ctypes.data, skw.nbytes)
    R.ref.call("Plaintext_SetParmsId", rskd, R.key_parms_id)
    rdec = R.decryptor(rsk)
    # relin keys -> reference
    cnt = u64()
    O.S.call("KSwitchKeys_GetKeyList", rlk, u64(0), C.byref(cnt), None)
    lst = (vp * cnt.value)()
    O.S.call("KSwitchKeys_GetKeyList", rlk, u64(0), C.byref(cnt), lst)
    words = []
    for h in lst:
        d = vp()
        O.S.call("PublicKey_Data", vp(h), C.byref(d))
        words.append(O.ct_words_key(d))
    rrlk = R.new_ksk({0: np.stack(words)})
    # encrypt with OUR encryptor (pk) and symmetric encryptor (sk)
    enc = vp()
    O.S.call("Encryptor_Create", O.ctx, pk, sk, C.byref(enc))
    rng = np.random.default_rng(8)
    m1 = rng.integers(0, t, size=n, dtype=np.uint64)
    m2 = rng.integers(0, t, size=16, dtype=np.uint64)
    c1, c2 = O._dst(), O._dst()
    O.S.call("Encryptor_Encrypt", enc, O.new_pt(m1), c1, None)
    O.S.call("Encryptor_EncryptSymmetric", enc, O.new_pt(m2), C.c_bool(False), c2, None)
    r1, r2 = R.new_ct(O.ct_words(c1)), R.new_ct(O.ct_words(c2))
    assert R.noise_budget(rdec, r1) > 20 and R.noise_budget(rdec, r2) > 20
    eq(R.pt_coeffs(R.decrypt(rdec, r1)), m1[: np.flatnonzero(m1)[-1] + 1], "reference decrypts our pk-encryption")
    eq(R.pt_coeffs(R.decrypt(rdec, r2)), m2[: np.flatnonzero(m2)[-1] + 1], "reference decrypts our sk-encryption")
    # reference evaluator with OUR relin keys: (c1 * c2) relinearized decrypts to the negacyclic product mod t
    rprod = R.relinearize(R.multiply(r1, r2), rrlk)
    assert R.noise_budget(rdec, rprod) > 0
    exp = np.zeros(n, dtype=object)
    for i, a in enumerate(m2):
        if a:
            shifted = np.concatenate([-(m1[n - i:].astype(object)), m1[: n - i].astype(object)]) if i else m1.astype(object)
            exp = (exp + int(a) * shifted) % t
    got = R.pt_coeffs(R.decrypt(rdec, rprod)).astype(object)
    full = np.zeros(n, dtype=object)
    full[: got.size] = got
    assert np.array_equal(full, exp % t), "product under our relinearization keys decrypts wrongly on the reference"
    # and entirely inside our layer
    odec = O.decryptor(skw)
    oprod = O.relinearize(O.multiply(c1, c2), rlk)
    eq(O.ct_words(oprod), R.ct_words(rprod), "our evaluator == reference evaluator on our keys")
    got2 = np.zeros(n, dtype=object)
    g = O.pt_coeffs(O.decrypt(odec, oprod)).astype(object)
    got2[: g.size] = g
    assert np.array_equal(got2, exp % t)
    if t % (2 * n) == 1:
        glk = vp()
        steps = (C.c_int * 2)(1, -2)
        O.S.call("KeyGenerator_CreateGaloisKeysFromSteps", kg, u64(2), steps, C.c_bool(False), C.byref(glk))
        be_ = R.batch_encoder()
        vals = rng.integers(0, t, size=n, dtype=np.uint64)
        cv = O._dst()
        O.S.call("Encryptor_Encrypt", enc, O.new_pt(R.pt_coeffs(R.batch_encode(be_, vals))), cv, None)
        rot = O.rotate_rows(cv, 1, glk)
        back = R.batch_decode(be_, R.decrypt(rdec, R.new_ct(O.ct_words(rot))))
        half = n // 2
        expect = np.concatenate([np.roll(vals[:half], -1), np.roll(vals[half:], -1)])
        eq(back, expect, "rotate_rows(1) with our Galois keys")


def chi_sq_dag(S, n, moduli, t, evaluations=2):
    """BASELINE config 4: the optimised chi-squared circuit (examples/chi_sq/src/main.rs:59-88) as the DAG
    sunscreen_runtime executes it — every Multiply followed by the Relinearize the compiler inserts
    (sunscreen_backend/src/transforms/insert_relinearizations.rs:17-62) — replayed through both C ABIs on the same
    fresh encryptions; every output ciphertext word must match, and the outputs decrypt to the plain computation."""
    R = refseal.RefContext(n, moduli, t)
    O = S.context(n, moduli, t)
    kg = R.keygen()
    sk, pk, rk = R.secret_key(kg), R.public_key(kg), R.relin_keys(kg)
    enc, dec = R.encryptor(pk), R.decryptor(sk)
    ork = O.new_ksk(R.ksk_words(rk))
    rng = np.random.default_rng(12)

    def circuit(E, rlk, n0, n1, n2):
        mul = lambda a, b: E.relinearize(E.multiply(a, b), rlk)
        x = E.add(E.add(n0, n0), n1)
        y = E.add(E.add(n2, n2), n1)
        a = mul(n0, n2)
        a = E.add(a, a)
        a = E.add(a, a)
        alpha = E.sub(a, mul(n1, n1))
        alpha = mul(alpha, alpha)
        b1 = mul(x, x)
        b1 = E.add(b1, b1)
        b2 = mul(x, y)
        b3 = mul(y, y)
        b3 = E.add(b3, b3)
        return alpha, b1, b2, b3

    for _ in range(evaluations):
        vals = [int(v) for v in rng.integers(1, 12, size=3)]
        rin = [R.encrypt(enc, R.new_pt(np.array([v], dtype=np.uint64))) for v in vals]
        oin = [O.new_ct(R.ct_words(h)) for h in rin]
        rout = circuit(R, rk, *rin)
        oout = circuit(O, ork, *oin)
        n0, n1, n2 = vals
        x, y = 2 * n0 + n1, 2 * n2 + n1
        expect = [(4 * n0 * n2 - n1 * n1) ** 2, 2 * x * x, x * y, 2 * y * y]
        for name, hr, ho, e in zip(("alpha", "b_1", "b_2", "b_3"), rout, oout, expect):
            eq(O.ct_words(ho), R.ct_words(hr), f"chi_sq output {name}")
            got = R.pt_coeffs(R.decrypt(dec, R.new_ct(O.ct_words(ho))))
            assert int(got[0]) == e % t and got.size == 1, (name, got[:4], e)


def rotate_multiply_plain_sweep(S, n, moduli, t, steps=(1, 2, 4, 64)):
    """BASELINE config 5: rotate_rows by powers of two followed by multiply_plain with a dense plaintext, with
    Galois keys for exactly those steps, against the reference word for word."""
    R = refseal.RefContext(n, moduli, t)
    O = S.context(n, moduli, t)
    inp = refseal.appendix_b_inputs(n, moduli, t)
    rng = np.random.default_rng(21)
    K = len(moduli)
    keys = {}
    for s in steps:
        elt = pow(3, s, 2 * n)
        key = np.empty((R.k, 2, K, n), dtype=np.uint64)
        for i in range(K):
            key[:, :, i, :] = rng.integers(0, moduli[i], size=(R.k, 2, n), dtype=np.uint64)
        keys[(elt - 1) // 2] = key
    rg, og = R.new_ksk(keys), O.new_ksk(keys)
    ra, oa = R.new_ct(inp["a"]), O.new_ct(inp["a"])
    rp, op = R.new_pt(inp["p"]), O.new_pt(inp["p"])
    for s in steps:
        rr, orr = R.rotate_rows(ra, s, rg), O.rotate_rows(oa, s, og)
        eq(O.ct_words(orr), R.ct_words(rr), f"rotate_rows({s})")
        eq(O.ct_words(O.multiply_plain(orr, op)), R.ct_words(R.multiply_plain(rr, rp)), f"multiply_plain after rotate_rows({s})")


def batch_encoder_parity(S, n, moduli, t):
    """BatchEncoder_Encode/Decode (slot permutation + negacyclic NTT mod t) against the reference."""
    R = refseal.RefContext(n, moduli, t)
    O = S.context(n, moduli, t)
    rbe = R.batch_encoder()
    obe = vp()
    O.S.call("BatchEncoder_Create", O.ctx, C.byref(obe))
    cnt = u64()
    O.S.call("BatchEncoder_GetSlotCount", obe, C.byref(cnt))
    assert cnt.value == n
    rng = np.random.default_rng(33)
    for size in (n, 7, 0):
        vals = rng.integers(0, t, size=size, dtype=np.uint64)
        rp = R.batch_encode(rbe, vals)
        op = vp()
        O.S.call("Plaintext_Create1", None, C.byref(op))
        O.S.call("BatchEncoder_Encode1", obe, u64(size), vals.ctypes.data_as(C.POINTER(u64)), op)
        eq(O.pt_coeffs(op), R.pt_coeffs(rp), f"BatchEncoder_Encode1 ({size} values)")
        out = np.zeros(n, dtype=np.uint64)
        c2 = u64(n)
        O.S.call("BatchEncoder_Decode1", obe, op, C.byref(c2), out.ctypes.data_as(C.POINTER(u64)), None)
        exp = np.zeros(n, dtype=np.uint64)
        exp[:size] = vals
        eq(out, exp, "BatchEncoder_Decode1 round trip")
    # signed variant
    sv = rng.integers(-(t // 2), t // 2, size=n, dtype=np.int64)
    op = vp()
    O.S.call("Plaintext_Create1", None, C.byref(op))
    O.S.call("BatchEncoder_Encode2", obe, u64(n), sv.ctypes.data_as(C.POINTER(C.c_int64)), op)
    rp = vp()
    R.ref.call("Plaintext_Create1", None, C.byref(rp))
    R.ref.call("BatchEncoder_Encode2", rbe, u64(n), sv.ctypes.data_as(C.POINTER(C.c_int64)), rp)
    eq(O.pt_coeffs(op), R.pt_coeffs(rp), "BatchEncoder_Encode2")
    so = np.zeros(n, dtype=np.int64)
    c2 = u64(n)
    O.S.call("BatchEncoder_Decode2", obe, op, C.byref(c2), so.ctypes.data_as(C.POINTER(C.c_int64)), None)
    assert np.array_equal(so, sv)
