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
