"""The `seal_fhe` crate's own BFV evaluator unit tests (seal_fhe/src/bfv_evaluator.rs:305-960) restated against the
Python mirror of the crate API (sunscreen_b200/seal_fhe.py): same parameters (n=8192, CoefficientModulus::create(8192,
[50,30,30,50,50]), PlainModulus::batching(8192,20)), same structure (run_bfv_test / make_vec), keys and encryptions made
by OUR KeyGenerator / Encryptor, results checked per SIMD lane after decryption."""
import random

from sunscreen_b200 import seal_fhe as s


def run_bfv_test(test, expand_mod_chain=False):
    params = (s.BfvEncryptionParametersBuilder().set_poly_modulus_degree(8192)
              .set_coefficient_modulus(s.CoefficientModulus.create(8192, [50, 30, 30, 50, 50]))
              .set_plain_modulus(s.PlainModulus.batching(8192, 20)).build())
    ctx = s.Context(params, expand_mod_chain, s.SecurityLevel.TC128)
    gen = s.KeyGenerator(ctx)
    encoder = s.BFVEncoder(ctx)
    public_key, secret_key = gen.create_public_key(), gen.secret_key()
    encryptor = s.Encryptor.with_public_and_secret_key(ctx, public_key, secret_key)
    decryptor = s.Decryptor(ctx, secret_key)
    evaluator = s.BFVEvaluator(ctx)
    test(decryptor, encoder, encryptor, evaluator, gen)


def make_vec(encoder, rng):
    return [rng.randrange(-100, 100) for _ in range(encoder.get_slot_count())]  # small values: products stay below t/2


def make_small_vec(encoder, rng):
    return [rng.randrange(-16, 16) for _ in range(encoder.get_slot_count())]


def all_tests():
    rng = random.Random(7)
    done = []

    def t(fn):
        run_bfv_test(fn)
        done.append(fn.__name__)

    def can_negate(dec, enc_, encr, ev, _):
        a = make_vec(enc_, rng)
        b = enc_.decode_signed(dec.decrypt(ev.negate(encr.encrypt(enc_.encode_signed(a)))))
        assert [-x for x in a] == b

    def can_negate_inplace(dec, enc_, encr, ev, _):
        a = make_vec(enc_, rng)
        c = encr.encrypt(enc_.encode_signed(a))
        ev.negate_inplace(c)
        assert [-x for x in a] == enc_.decode_signed(dec.decrypt(c))

    def can_add_sub(dec, enc_, encr, ev, _):
        a, b = make_vec(enc_, rng), make_vec(enc_, rng)
        ca, cb = encr.encrypt(enc_.encode_signed(a)), encr.encrypt(enc_.encode_signed(b))
        assert enc_.decode_signed(dec.decrypt(ev.add(ca, cb))) == [x + y for x, y in zip(a, b)]
        assert enc_.decode_signed(dec.decrypt(ev.sub(ca, cb))) == [x - y for x, y in zip(a, b)]
        ev.add_inplace(ca, cb)
        assert enc_.decode_signed(dec.decrypt(ca)) == [x + y for x, y in zip(a, b)]

    def can_add_many(dec, enc_, encr, ev, _):
        vs = [make_vec(enc_, rng) for _ in range(4)]
        cs = [encr.encrypt(enc_.encode_signed(v)) for v in vs]
        assert enc_.decode_signed(dec.decrypt(ev.add_many(cs))) == [sum(col) for col in zip(*vs)]

    def can_multiply_and_relinearize(dec, enc_, encr, ev, gen):
        a, b = make_vec(enc_, rng), make_vec(enc_, rng)
        ca, cb = encr.encrypt(enc_.encode_signed(a)), encr.encrypt(enc_.encode_signed(b))
        rk = gen.create_relinearization_keys()
        prod = ev.multiply(ca, cb)
        assert prod.num_polynomials() == 3
        before = dec.invariant_noise_budget(prod)
        rel = ev.relinearize(prod, rk)
        assert rel.num_polynomials() == 2
        # relinearization consumes (almost) no budget (seal_fhe/tests/assumptions.rs:139-194)
        assert before - dec.invariant_noise_budget(rel) <= 1
        assert enc_.decode_signed(dec.decrypt(rel)) == [x * y for x, y in zip(a, b)]
        ev.multiply_inplace(ca, cb)
        ev.relinearize_inplace(ca, rk)
        assert enc_.decode_signed(dec.decrypt(ca)) == [x * y for x, y in zip(a, b)]

    def can_square(dec, enc_, encr, ev, _):
        a = make_vec(enc_, rng)
        assert enc_.decode_signed(dec.decrypt(ev.square(encr.encrypt(enc_.encode_signed(a))))) == [x * x for x in a]

    def can_multiply_many_and_exponentiate(dec, enc_, encr, ev, gen):
        vs = [make_small_vec(enc_, rng) for _ in range(3)]
        cs = [encr.encrypt(enc_.encode_signed(v)) for v in vs]
        rk = gen.create_relinearization_keys()
        assert enc_.decode_signed(dec.decrypt(ev.multiply_many(cs, rk))) == [x * y * z for x, y, z in zip(*vs)]
        assert enc_.decode_signed(dec.decrypt(ev.exponentiate(cs[0], 3, rk))) == [x ** 3 for x in vs[0]]

    def can_plain_ops(dec, enc_, encr, ev, _):
        a, b = make_vec(enc_, rng), make_vec(enc_, rng)
        ca, pb = encr.encrypt(enc_.encode_signed(a)), enc_.encode_signed(b)
        assert enc_.decode_signed(dec.decrypt(ev.add_plain(ca, pb))) == [x + y for x, y in zip(a, b)]
        assert enc_.decode_signed(dec.decrypt(ev.sub_plain(ca, pb))) == [x - y for x, y in zip(a, b)]
        assert enc_.decode_signed(dec.decrypt(ev.multiply_plain(ca, pb))) == [x * y for x, y in zip(a, b)]

    def can_rotate(dec, enc_, encr, ev, gen):
        a = make_vec(enc_, rng)
        ca = encr.encrypt(enc_.encode_signed(a))
        gk = gen.create_galois_keys()
        half = len(a) // 2
        rot = lambda v, k: v[k % half:half] + v[:k % half] + v[half + k % half:] + v[half:half + k % half]
        assert enc_.decode_signed(dec.decrypt(ev.rotate_rows(ca, -1, gk))) == rot(a, -1)   # bfv_evaluator.rs: can_rotate_rows
        assert enc_.decode_signed(dec.decrypt(ev.rotate_rows(ca, 5, gk))) == rot(a, 5)     # 5 = 4 + 1: NAF path
        assert enc_.decode_signed(dec.decrypt(ev.rotate_columns(ca, gk))) == a[half:] + a[:half]

    def can_mod_switch(dec, enc_, encr, ev, _):
        a = make_vec(enc_, rng)
        c = ev.mod_switch_to_next(encr.encrypt(enc_.encode_signed(a)))
        assert c.coeff_modulus_size() == 3
        assert enc_.decode_signed(dec.decrypt(c)) == a

    def cannot_mod_switch_without_chain(dec, enc_, encr, ev, _):
        # Context::new(.., expand_mod_chain = false, ..): the chain ends at the first data level (S/context.cpp:478-497)
        try:
            ev.mod_switch_to_next(encr.encrypt(enc_.encode_signed(make_vec(enc_, rng))))
        except s.Error as e:
            assert "InvalidArgument" in str(e)
        else:
            raise AssertionError("mod_switch_to_next must fail at the end of the chain")

    def symmetric_encryption_roundtrip(dec, enc_, encr, ev, _):
        a = make_vec(enc_, rng)
        assert enc_.decode_signed(dec.decrypt(encr.encrypt_symmetric(enc_.encode_signed(a)))) == a

    for fn in (can_negate, can_negate_inplace, can_add_sub, can_add_many, can_multiply_and_relinearize, can_square,
               can_multiply_many_and_exponentiate, can_plain_ops, can_rotate, cannot_mod_switch_without_chain,
               symmetric_encryption_roundtrip):
        t(fn)
    run_bfv_test(can_mod_switch, expand_mod_chain=True)
    done.append("can_mod_switch")
    return done


def serialization_and_components_tests():
    """plaintext_ciphertext.rs / key_generator.rs / encryptor_decryptor.rs / poly_array.rs unit tests, restated:
    can_{save,load}_* round trips through as_bytes / from_bytes, encrypt_deterministic is repeatable and decrypts,
    the exported components have the crate's shapes."""
    rng = random.Random(11)
    done = []

    def roundtrips(dec, enc_, encr, ev, gen):
        ctx = gen.ctx
        a = make_vec(enc_, rng)
        pt = enc_.encode_signed(a)
        ct = encr.encrypt(pt)
        for mode in (s.CompressionType.ZSTD, s.CompressionType.ZLIB, s.CompressionType.NONE):
            pt2 = s.Plaintext.from_bytes(ctx, pt.as_bytes(mode))
            assert enc_.decode_signed(pt2) == a
            ct2 = s.Ciphertext.from_bytes(ctx, ct.as_bytes(mode))
            assert ct2.as_bytes(s.CompressionType.NONE) == ct.as_bytes(s.CompressionType.NONE)
            assert enc_.decode_signed(dec.decrypt(ct2)) == a
        sk2 = s.SecretKey.from_bytes(ctx, gen.secret_key().as_bytes())
        assert enc_.decode_signed(s.Decryptor(ctx, sk2).decrypt(ct)) == a
        pk2 = s.PublicKey.from_bytes(ctx, gen.create_public_key().as_bytes())
        assert enc_.decode_signed(dec.decrypt(s.Encryptor.with_public_key(ctx, pk2).encrypt(pt))) == a
        rk2 = s.RelinearizationKeys.from_bytes(ctx, gen.create_relinearization_keys().as_bytes())
        small = make_small_vec(enc_, rng)
        cs = encr.encrypt(enc_.encode_signed(small))
        assert enc_.decode_signed(dec.decrypt(ev.relinearize(ev.multiply(cs, cs), rk2))) == [x * x for x in small]
        gk2 = s.GaloisKeys.from_bytes(ctx, gen.create_galois_keys().as_bytes())
        half = len(a) // 2
        assert enc_.decode_signed(dec.decrypt(ev.rotate_columns(ct, gk2))) == a[half:] + a[:half]
        try:
            s.Ciphertext.from_bytes(ctx, ct.as_bytes()[:-3])
        except s.Error:
            pass
        else:
            raise AssertionError("a truncated ciphertext must not load")

    def deterministic(dec, enc_, encr, ev, gen):
        a = make_vec(enc_, rng)
        pt = enc_.encode_signed(a)
        seed = [rng.getrandbits(63) for _ in range(8)]
        c1, c2 = encr.encrypt_deterministic(pt, seed), encr.encrypt_deterministic(pt, seed)
        assert c1.as_bytes(s.CompressionType.NONE) == c2.as_bytes(s.CompressionType.NONE)
        assert encr.encrypt(pt).as_bytes(s.CompressionType.NONE) != c1.as_bytes(s.CompressionType.NONE)
        assert enc_.decode_signed(dec.decrypt(c1)) == a
        ct, u, e, r = encr.encrypt_return_components(pt, False, seed)
        assert ct.as_bytes(s.CompressionType.NONE) == c1.as_bytes(s.CompressionType.NONE)
        assert (u.num_polynomials(), e.num_polynomials()) == (1, 2) and u.coeff_modulus_size() == 5 and r.len() == pt.len()
        assert u.poly_modulus_degree() == 8192 and set(u.as_u64_slice()[:8192]) <= {0, 1, gen.ctx.params.get_coefficient_modulus()[0].value() - 1}
        assert dec.invariant_noise_budget(ct) > 0 and 0.0 < dec.invariant_noise(ct) < 2.0 ** -20

    def polynomial_arrays(dec, enc_, encr, ev, gen):
        ctx = gen.ctx
        ct = encr.encrypt(enc_.encode_signed(make_vec(enc_, rng)))
        pa = s.PolynomialArray.new_from_ciphertext(ctx, ct)
        assert (pa.num_polynomials(), pa.coeff_modulus_size(), pa.poly_modulus_degree()) == (2, 4, 8192) and pa.is_rns()
        words = pa.as_u64_slice()
        assert words[:8] == [ct.get_data(i) for i in range(8)]
        pa.to_multiprecision()
        assert pa.is_multiprecision()
        pa.to_rns()
        assert pa.as_u64_slice() == words
        pk = s.PolynomialArray.new_from_public_key(ctx, gen.create_public_key())
        assert (pk.num_polynomials(), pk.coeff_modulus_size()) == (2, 4)
        assert pk.drop().coeff_modulus_size() == 3 and pk.clone().as_u64_slice() == pk.as_u64_slice()
        sk = s.PolynomialArray.new_from_secret_key(ctx, gen.secret_key())
        assert sk.num_polynomials() == 1 and set(sk.as_u64_slice()[:8192]) <= {0, 1, gen.ctx.params.get_coefficient_modulus()[0].value() - 1}
        assert not s.PolynomialArray().is_reserved()

    def hex_strings(dec, enc_, encr, ev, gen):
        p = s.Plaintext.from_hex_string("1234x^2 + 4321")
        assert (p.len(), p.get_coefficient(0), p.get_coefficient(1), p.get_coefficient(2)) == (3, 0x4321, 0, 0x1234)

    for fn in (roundtrips, deterministic, polynomial_arrays, hex_strings):
        run_bfv_test(fn)
        done.append(fn.__name__)
    return done


def lane_overflow_assumption():
    """seal_fhe/tests/assumptions.rs:5-34: lanes wrap modulo the plain modulus (default parameters, n=8192, t=114689?)
    restated: with t = PlainModulus::batching(8192, 17), 300*400 wraps to 120000 mod t in every lane."""
    params = (s.BfvEncryptionParametersBuilder().set_poly_modulus_degree(8192)
              .set_coefficient_modulus(s.CoefficientModulus.bfv_default(8192, s.SecurityLevel.TC128))
              .set_plain_modulus(s.PlainModulus.batching(8192, 17)).build())
    t = params.get_plain_modulus().value()
    ctx = s.Context(params, True, s.SecurityLevel.TC128)
    gen = s.KeyGenerator(ctx)
    enc_ = s.BFVEncoder(ctx)
    encr = s.Encryptor.with_public_key(ctx, gen.create_public_key())
    dec = s.Decryptor(ctx, gen.secret_key())
    ev = s.BFVEvaluator(ctx)
    n = enc_.get_slot_count()
    ca, cb = encr.encrypt(enc_.encode_unsigned([300] * n)), encr.encrypt(enc_.encode_unsigned([400] * n))
    out = enc_.decode_unsigned(dec.decrypt(ev.multiply(ca, cb)))
    assert out == [120000 % t] * n and 120000 > t
