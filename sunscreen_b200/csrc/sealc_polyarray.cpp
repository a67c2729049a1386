// sealc_polyarray.cpp — PolynomialArray_* of the reference's C export layer (S/c/polyarray.cpp:14-253 over
// S/polyarray.{h,cpp}): the container seal_fhe uses to hand the encryption components (u, e) and key polynomials
// to the zero-knowledge proof code.  Transforms out of NTT form run on the GPU through layer 1; the RNS <-> multi-
// precision conversions are host integer arithmetic (S/util/rns.cpp:288-412).
//
// Behaviours that look odd are the reference's and are kept on purpose (a drop-in must agree with it):
//   * CreateFromSecretKey keeps ONE "polynomial" of K*n words over the base {t} and inverse-transforms only the
//     first k (data-level) residues (S/polyarray.cpp:88-116);
//   * Drop copies the first poly_size*n*(k-1) words linearly (S/polyarray.cpp:186-209), which is a per-polynomial
//     drop of the last residue only when poly_size == 1.
#include "../../include/b200_sealc.h"
#include "sealc_types.h"

namespace
{
using namespace b200c;
typedef unsigned __int128 u128;

std::vector<u64> first_level_moduli(Context_ *c)
{
    return std::vector<u64>(c->parms.coeff.begin(), c->parms.coeff.begin() + c->level_k[c->first_level]);
}

// inverse NTT of `polys` polynomials of k = first-level residues each, laid out [poly][k][n] in `w`
void intt_first_level(Context_ *c, u64 *w, size_t polys)
{
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t words = polys * (size_t)c->level_k[c->first_level] * c->parms.n;
    DevBuf d(c, words);
    dev_check(b200_memcpy_h2d(c->dev, d.p, w, words * 8, nullptr));
    dev_check(b200_ntt_inverse(c->dev, c->first_level, d.p, polys, nullptr));
    dev_check(b200_memcpy_d2h(c->dev, w, d.p, words * 8, nullptr));
    dev_check(b200_stream_synchronize(c->dev, nullptr));
}

// multi-word helpers (little-endian words)
void mp_mul_word_add_mod(std::vector<u64> &acc, const std::vector<u64> &a, u64 w, const std::vector<u64> &mod)
{
    // acc = (acc + a * w) mod `mod`, all operands < mod, size = mod.size()
    const size_t W = mod.size();
    std::vector<u64> prod(W + 1, 0);
    u64 carry = 0;
    for (size_t i = 0; i < W; i++)
    {
        u128 m = (u128)a[i] * w + carry;
        prod[i] = (u64)m;
        carry = (u64)(m >> 64);
    }
    prod[W] = carry;
    // reduce prod (W+1 words) modulo mod by schoolbook shift-subtract on bits (W <= 64, host utility path)
    std::vector<u64> rem(W + 1, 0);
    for (size_t bit = (W + 1) * 64; bit-- > 0;)
    {
        // rem = (rem << 1) | bit
        u64 c = (prod[bit >> 6] >> (bit & 63)) & 1;
        for (size_t i = 0; i <= W; i++)
        {
            u64 nc = rem[i] >> 63;
            rem[i] = (rem[i] << 1) | c;
            c = nc;
        }
        // if rem >= mod: rem -= mod
        bool ge = rem[W] != 0;
        if (!ge)
        {
            ge = true;
            for (size_t i = W; i-- > 0;)
                if (rem[i] != mod[i])
                {
                    ge = rem[i] > mod[i];
                    break;
                }
        }
        if (ge)
        {
            u64 borrow = 0;
            for (size_t i = 0; i <= W; i++)
            {
                const u64 m = i < W ? mod[i] : 0;
                const u64 t = rem[i] - m - borrow;
                borrow = (rem[i] < m + borrow) || (m + borrow < m) ? 1 : 0;
                rem[i] = t;
            }
        }
    }
    // acc = (acc + rem) mod mod
    u64 carry2 = 0;
    for (size_t i = 0; i < W; i++)
    {
        u128 s = (u128)acc[i] + rem[i] + carry2;
        acc[i] = (u64)s;
        carry2 = (u64)(s >> 64);
    }
    bool ge = carry2 != 0;
    if (!ge)
    {
        ge = true;
        for (size_t i = W; i-- > 0;)
            if (acc[i] != mod[i])
            {
                ge = acc[i] > mod[i];
                break;
            }
    }
    if (ge)
    {
        u64 borrow = 0;
        for (size_t i = 0; i < W; i++)
        {
            const u64 t = acc[i] - mod[i] - borrow;
            borrow = (acc[i] < mod[i] + borrow) || (mod[i] + borrow < mod[i]) ? 1 : 0;
            acc[i] = t;
        }
    }
}

struct Crt
{
    std::vector<u64> q;
    std::vector<u64> prod;               // Q, k words
    std::vector<std::vector<u64>> punct; // Q/q_i, k words
    std::vector<u64> inv_punct;          // (Q/q_i)^-1 mod q_i
    explicit Crt(const std::vector<u64> &moduli) : q(moduli)
    {
        const size_t k = q.size();
        auto mul_word = [&](std::vector<u64> &a, u64 w) {
            u64 carry = 0;
            for (size_t i = 0; i < a.size(); i++)
            {
                u128 m = (u128)a[i] * w + carry;
                a[i] = (u64)m;
                carry = (u64)(m >> 64);
            }
        };
        prod.assign(k, 0);
        prod[0] = 1;
        for (u64 m : q)
            mul_word(prod, m);
        punct.resize(k);
        inv_punct.resize(k);
        for (size_t i = 0; i < k; i++)
        {
            punct[i].assign(k, 0);
            punct[i][0] = 1;
            u64 r = 1 % q[i];
            for (size_t j = 0; j < k; j++)
                if (j != i)
                {
                    mul_word(punct[i], q[j]);
                    r = (u64)((u128)r * (q[j] % q[i]) % q[i]);
                }
            inv_punct[i] = b200::inv_mod(r, q[i]);
        }
    }
};

// RNSBase::compose_array (S/util/rns.cpp:366-412): [k][count] residues -> [count][k] words, in place
void compose_array(u64 *value, size_t count, const Crt &crt)
{
    const size_t k = crt.q.size();
    if (k <= 1)
        return;
    std::vector<u64> tmp(value, value + count * k);
    std::vector<u64> acc(k);
    for (size_t i = 0; i < count; i++)
    {
        std::fill(acc.begin(), acc.end(), 0);
        for (size_t j = 0; j < k; j++)
        {
            const u64 t = (u64)((u128)tmp[j * count + i] * crt.inv_punct[j] % crt.q[j]);
            mp_mul_word_add_mod(acc, crt.punct[j], t, crt.prod);
        }
        std::copy(acc.begin(), acc.end(), value + i * k);
    }
}
// RNSBase::decompose_array (S/util/rns.cpp:288-325): [count][k] words -> [k][count] residues, in place
void decompose_array(u64 *value, size_t count, const Crt &crt)
{
    const size_t k = crt.q.size();
    if (k <= 1)
        return;
    std::vector<u64> tmp(value, value + count * k);
    for (size_t j = 0; j < k; j++)
    {
        const u64 q = crt.q[j];
        for (size_t i = 0; i < count; i++)
        {
            u64 r = 0;
            for (size_t w = k; w-- > 0;)
                r = (u64)((((u128)r << 64) | tmp[i * k + w]) % q);
            value[j * count + i] = r;
        }
    }
}

} // namespace

extern "C" {

long PolynomialArray_Create(void *, void **out)
{
    NULLRET(out);
    *out = new PolynomialArray_();
    return S_OK_;
}
long PolynomialArray_CreateFromCiphertext(void *, void *context, void *ciphertext, void **out)
{
    NULLRET(context);
    NULLRET(ciphertext);
    NULLRET(out);
    auto *c = (Context_ *)context;
    auto *ct = (Ciphertext_ *)ciphertext;
    auto *pa = new PolynomialArray_();
    long hr = guard([&] {
        if (!c->parameters_set)
            throw InvalidArg("encryption parameters are not set correctly");
        ct->sync_host();
        pa->reserve((size_t)ct->size, (size_t)ct->n, first_level_moduli(c));
        const size_t stride = (size_t)(ct->n * ct->k);
        if (ct->size && (ct->size - 1) * stride + pa->poly_len() > ct->host.size())
            throw InvalidArg("ciphertext is not at the first data level");
        for (size_t i = 0; i < ct->size; i++)
            pa->insert(i, ct->host.data() + i * stride);
        if (ct->is_ntt_form && ct->size)
            intt_first_level(c, pa->data.data(), (size_t)ct->size);
    });
    if (hr)
    {
        delete pa;
        return hr;
    }
    *out = pa;
    return S_OK_;
}
long PolynomialArray_CreateFromPublicKey(void *, void *context, void *public_key, void **out)
{
    NULLRET(context);
    NULLRET(public_key);
    NULLRET(out);
    auto *c = (Context_ *)context;
    auto &ct = ((PublicKey_ *)public_key)->data;
    auto *pa = new PolynomialArray_();
    long hr = guard([&] {
        if (!c->parameters_set)
            throw InvalidArg("encryption parameters are not set correctly");
        ct.sync_host();
        pa->reserve((size_t)ct.size, (size_t)ct.n, first_level_moduli(c));
        const size_t stride = (size_t)(ct.n * ct.k); // key level: the special prime's residue is skipped
        if (ct.size && (ct.size - 1) * stride + pa->poly_len() > ct.host.size())
            throw InvalidArg("public key is not valid for encryption parameters");
        for (size_t i = 0; i < ct.size; i++)
            pa->insert(i, ct.host.data() + i * stride);
        if (ct.is_ntt_form && ct.size)
            intt_first_level(c, pa->data.data(), (size_t)ct.size);
    });
    if (hr)
    {
        delete pa;
        return hr;
    }
    *out = pa;
    return S_OK_;
}
long PolynomialArray_CreateFromSecretKey(void *, void *context, void *secret_key, void **out)
{
    NULLRET(context);
    NULLRET(secret_key);
    NULLRET(out);
    auto *c = (Context_ *)context;
    auto &pt = ((SecretKey_ *)secret_key)->data;
    auto *pa = new PolynomialArray_();
    long hr = guard([&] {
        if (!c->parameters_set)
            throw InvalidArg("encryption parameters are not set correctly");
        pa->reserve(1, pt.coeffs.size(), std::vector<u64>{ c->parms.plain });
        pa->insert(0, pt.coeffs.data());
        const size_t k = (size_t)c->level_k[c->first_level];
        if (pt.parms_id != kZeroId)
        {
            if (pt.coeffs.size() < k * c->parms.n)
                throw InvalidArg("secret key is not valid for encryption parameters");
            intt_first_level(c, pa->data.data(), 1);
        }
    });
    if (hr)
    {
        delete pa;
        return hr;
    }
    *out = pa;
    return S_OK_;
}
long PolynomialArray_Copy(void *copy, void **out)
{
    NULLRET(copy);
    NULLRET(out);
    // the reference's copy constructor re-reserves and re-inserts the filled polynomials, always as an RNS array
    // (S/polyarray.cpp:118-136)
    auto *src = (PolynomialArray_ *)copy;
    auto *pa = new PolynomialArray_();
    pa->reserve(src->poly_size, src->coeff_size, src->moduli);
    for (size_t i = 0; i < src->poly_size; i++)
        if (src->filled[i])
            pa->insert(i, src->data.data() + i * src->poly_len());
    *out = pa;
    return S_OK_;
}
long PolynomialArray_Destroy(void *p)
{
    NULLRET(p);
    delete (PolynomialArray_ *)p;
    return S_OK_;
}
long PolynomialArray_IsReserved(void *p, bool *r)
{
    NULLRET(p);
    NULLRET(r);
    *r = ((PolynomialArray_ *)p)->reserved;
    return S_OK_;
}
long PolynomialArray_IsRns(void *p, bool *r)
{
    NULLRET(p);
    NULLRET(r);
    *r = ((PolynomialArray_ *)p)->is_rns;
    return S_OK_;
}
long PolynomialArray_IsMultiprecision(void *p, bool *r)
{
    NULLRET(p);
    NULLRET(r);
    *r = !((PolynomialArray_ *)p)->is_rns;
    return S_OK_;
}
long PolynomialArray_ToRns(void *p)
{
    NULLRET(p);
    auto *pa = (PolynomialArray_ *)p;
    if (pa->is_rns)
        return S_OK_;
    return guard([&] {
        Crt crt(pa->moduli);
        for (size_t i = 0; i < pa->poly_size; i++)
            decompose_array(pa->data.data() + i * pa->poly_len(), pa->coeff_size, crt);
        pa->is_rns = true;
    });
}
long PolynomialArray_ToMultiprecision(void *p)
{
    NULLRET(p);
    auto *pa = (PolynomialArray_ *)p;
    if (!pa->is_rns)
        return S_OK_;
    return guard([&] {
        Crt crt(pa->moduli);
        for (size_t i = 0; i < pa->poly_size; i++)
            compose_array(pa->data.data() + i * pa->poly_len(), pa->coeff_size, crt);
        pa->is_rns = false;
    });
}
long PolynomialArray_GetPolynomial(void *p, uint64_t poly_index, uint64_t *data)
{
    NULLRET(p);
    NULLRET(data);
    auto *pa = (PolynomialArray_ *)p;
    if (poly_index >= pa->poly_size)
        return COR_E_INVALIDOPERATION_; // the reference throws logic_error here, which its C layer does not catch
    *data = pa->data[poly_index * pa->poly_len()]; // first word only, as in S/c/polyarray.cpp:177-190
    return S_OK_;
}
long PolynomialArray_ExportSize(void *p, uint64_t *size)
{
    NULLRET(p);
    NULLRET(size);
    *size = ((PolynomialArray_ *)p)->data.size();
    return S_OK_;
}
long PolynomialArray_PerformExport(void *p, uint64_t *data)
{
    NULLRET(p);
    NULLRET(data);
    auto *pa = (PolynomialArray_ *)p;
    std::copy(pa->data.begin(), pa->data.end(), data);
    return S_OK_;
}
long PolynomialArray_PolySize(void *p, uint64_t *size)
{
    NULLRET(p);
    NULLRET(size);
    *size = ((PolynomialArray_ *)p)->poly_size;
    return S_OK_;
}
long PolynomialArray_PolyModulusDegree(void *p, uint64_t *size)
{
    NULLRET(p);
    NULLRET(size);
    *size = ((PolynomialArray_ *)p)->coeff_size;
    return S_OK_;
}
long PolynomialArray_CoeffModulusSize(void *p, uint64_t *size)
{
    NULLRET(p);
    NULLRET(size);
    *size = ((PolynomialArray_ *)p)->moduli.size();
    return S_OK_;
}
long PolynomialArray_Drop(void *p, void **out)
{
    NULLRET(p);
    NULLRET(out);
    auto *src = (PolynomialArray_ *)p;
    auto *pa = new PolynomialArray_();
    long hr = guard([&] {
        if (src->moduli.size() < 2)
            throw LogicErr("cannot drop from base of size 1"); // RNSBase::drop (S/util/rns.h)
        std::vector<u64> lower(src->moduli.begin(), src->moduli.end() - 1);
        pa->reserve(src->poly_size, src->coeff_size, lower);
        std::copy_n(src->data.begin(), pa->data.size(), pa->data.begin());
        pa->filled = src->filled;
    });
    if (hr)
    {
        delete pa;
        return hr;
    }
    *out = pa;
    return S_OK_;
}

} // extern "C"
