// oracle/ref_shim.cpp — TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" hooks compiled INTO oracle/_ref/libsealc_ref.so next to the unmodified
// reference sources (they #include the reference's own headers from /root/reference; no
// reference code is copied).  They exist because the reference's C export layer only offers
// word-at-a-time data access (Ciphertext_GetDataAt1 / Ciphertext_SetDataAt, S/c/ciphertext.h),
// which is far too slow for differential tests over 10^5..10^7 words, and because the parity
// tests want to call a few util-level reference functions directly:
//   ntt_negacyclic_harvey / inverse_ntt_negacyclic_harvey   (S/util/ntt.cpp:393-474)
//   RNSTool::fastbconv_m_tilde + sm_mrq                     (S/util/rns.cpp:991-1143)
//   RNSTool::fast_floor + fastbconv_sk                      (S/util/rns.cpp:915-1096)
// plus a multi-threaded timing loop over Evaluator::multiply + relinearize_inplace that
// bench.py reports as the CPU baseline (`cpu_baseline.kind = "reference"`).
//
// Handles are the same void* objects the reference C API hands out (S/c/utilities.h FromVoid).
#include "seal/seal.h"
#include "seal/util/ntt.h"
#include "seal/util/rns.h"
#include "seal/util/polyarithsmallmod.h"
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

using namespace seal;
using namespace seal::util;

extern "C" {

uint64_t *refshim_ct_data(void *ct) { return reinterpret_cast<Ciphertext *>(ct)->data(); }
uint64_t refshim_ct_words(void *ct)
{
    auto *c = reinterpret_cast<Ciphertext *>(ct);
    return c->size() * c->coeff_modulus_size() * c->poly_modulus_degree();
}
int refshim_ct_resize(void *ct, void *ctx, const uint64_t *parms_id, uint64_t size, int is_ntt)
{
    try
    {
        parms_id_type pid;
        std::copy_n(parms_id, pid.size(), pid.begin());
        auto *c = reinterpret_cast<Ciphertext *>(ct);
        c->resize(*reinterpret_cast<SEALContext *>(ctx), pid, size);
        c->is_ntt_form() = is_ntt != 0;
        return 0;
    }
    catch (...) { return -1; }
}
uint64_t *refshim_pt_data(void *pt) { return reinterpret_cast<Plaintext *>(pt)->data(); }
uint64_t refshim_pt_coeff_count(void *pt) { return reinterpret_cast<Plaintext *>(pt)->coeff_count(); }

// Number of key lists (outer vector) and whether list `index` is populated.
uint64_t refshim_ksk_outer_size(void *keys) { return reinterpret_cast<KSwitchKeys *>(keys)->data().size(); }
uint64_t refshim_ksk_inner_size(void *keys, uint64_t index)
{
    auto &d = reinterpret_cast<KSwitchKeys *>(keys)->data();
    return index < d.size() ? d[index].size() : 0;
}
// Pointer to the raw words of key[index][j] (a size-2 key-level ciphertext, NTT form).
uint64_t *refshim_ksk_data(void *keys, uint64_t index, uint64_t j)
{
    return reinterpret_cast<KSwitchKeys *>(keys)->data()[index][j].data().data();
}
// Make key list `index` hold `decomp` size-2 key-level ciphertexts (contents undefined until written).
int refshim_ksk_alloc(void *keys, void *ctx, uint64_t index, uint64_t decomp)
{
    try
    {
        auto *k = reinterpret_cast<KSwitchKeys *>(keys);
        auto &context = *reinterpret_cast<SEALContext *>(ctx);
        if (k->data().size() <= index)
            k->data().resize(index + 1);
        k->data()[index].clear();
        for (uint64_t j = 0; j < decomp; j++)
        {
            PublicKey pk;
            pk.data().resize(context, context.key_parms_id(), 2);
            pk.data().is_ntt_form() = true;
            pk.parms_id() = context.key_parms_id();
            k->data()[index].push_back(std::move(pk));
        }
        k->parms_id() = context.key_parms_id();
        return 0;
    }
    catch (...) { return -1; }
}

static std::mutex g_ntt_mutex;
static std::map<std::pair<uint64_t, int>, std::unique_ptr<NTTTables>> g_ntt_cache;
static const NTTTables *get_tables(uint64_t modulus, int logn)
{
    std::lock_guard<std::mutex> lk(g_ntt_mutex);
    auto key = std::make_pair(modulus, logn);
    auto it = g_ntt_cache.find(key);
    if (it == g_ntt_cache.end())
        it = g_ntt_cache.emplace(key, std::make_unique<NTTTables>(logn, Modulus(modulus))).first;
    return it->second.get();
}
// In-place transforms of `count` consecutive size-2^logn polynomials modulo `modulus`; fully reduced output.
int refshim_ntt_forward(uint64_t modulus, int logn, uint64_t *data, uint64_t count)
{
    try
    {
        auto *t = get_tables(modulus, logn);
        for (uint64_t i = 0; i < count; i++)
            ntt_negacyclic_harvey(CoeffIter(data + (i << logn)), *t);
        return 0;
    }
    catch (...) { return -1; }
}
int refshim_ntt_inverse(uint64_t modulus, int logn, uint64_t *data, uint64_t count)
{
    try
    {
        auto *t = get_tables(modulus, logn);
        for (uint64_t i = 0; i < count; i++)
            inverse_ntt_negacyclic_harvey(CoeffIter(data + (i << logn)), *t);
        return 0;
    }
    catch (...) { return -1; }
}
uint64_t refshim_ntt_root(uint64_t modulus, int logn) { return get_tables(modulus, logn)->get_root(); }

// --- context constants (for pinning the host precompute) -------------------------------------
static std::shared_ptr<const SEALContext::ContextData> ctxdata(void *ctx, int key_level)
{
    auto &context = *reinterpret_cast<SEALContext *>(ctx);
    return key_level ? context.key_context_data() : context.first_context_data();
}
// out: [|B|, |Bsk|, m_sk, gamma, t, Bsk primes...]
int refshim_rns_info(void *ctx, int key_level, uint64_t *out, uint64_t cap)
{
    auto cd = ctxdata(ctx, key_level);
    auto *rt = cd->rns_tool();
    size_t nb = rt->base_B()->size(), nbsk = rt->base_Bsk()->size();
    if (cap < 5 + nbsk)
        return -1;
    out[0] = nb;
    out[1] = nbsk;
    out[2] = rt->m_sk().value();
    out[3] = rt->gamma().value();
    out[4] = rt->t().value();
    for (size_t i = 0; i < nbsk; i++)
        out[5 + i] = (*rt->base_Bsk())[i].value();
    return 0;
}
// out: [delta_mod_q_i (k), upper_half_increment_i (k), plain_upper_half_increment_i (k), plain_upper_half_threshold]
int refshim_plain_info(void *ctx, int key_level, uint64_t *out, uint64_t cap)
{
    auto cd = ctxdata(ctx, key_level);
    size_t k = cd->parms().coeff_modulus().size();
    if (cap < 3 * k + 1)
        return -1;
    for (size_t i = 0; i < k; i++)
    {
        out[i] = cd->coeff_div_plain_modulus()[i].operand;
        out[k + i] = cd->upper_half_increment()[i];
        out[2 * k + i] = cd->plain_upper_half_increment()[i];
    }
    out[3 * k] = cd->plain_upper_half_threshold();
    return 0;
}

// BEHZ steps 1-2 on one base-q polynomial: in = k*n words, out = |Bsk|*n words (S/evaluator.cpp:476-480).
int refshim_behz_lift(void *ctx, const uint64_t *in, uint64_t *out)
{
    try
    {
        auto cd = ctxdata(ctx, 0);
        auto *rt = cd->rns_tool();
        size_t n = cd->parms().poly_modulus_degree();
        size_t nbskm = rt->base_Bsk_m_tilde()->size();
        auto pool = MemoryManager::GetPool();
        std::vector<uint64_t> tmp(nbskm * n);
        rt->fastbconv_m_tilde(ConstRNSIter(in, n), RNSIter(tmp.data(), n), pool);
        rt->sm_mrq(ConstRNSIter(tmp.data(), n), RNSIter(out, n), pool);
        return 0;
    }
    catch (...) { return -1; }
}
// BEHZ steps 7-8: in = (k+|Bsk|)*n words already multiplied by t; out = k*n words (S/evaluator.cpp:560-565).
int refshim_behz_floor_sk(void *ctx, const uint64_t *in, uint64_t *out)
{
    try
    {
        auto cd = ctxdata(ctx, 0);
        auto *rt = cd->rns_tool();
        size_t n = cd->parms().poly_modulus_degree();
        size_t nbsk = rt->base_Bsk()->size();
        auto pool = MemoryManager::GetPool();
        std::vector<uint64_t> tmp(nbsk * n);
        rt->fast_floor(ConstRNSIter(in, n), RNSIter(tmp.data(), n), pool);
        rt->fastbconv_sk(ConstRNSIter(tmp.data(), n), RNSIter(out, n), pool);
        return 0;
    }
    catch (...) { return -1; }
}

// --- full-size parity: `count` independent (a, b) pairs of size-2 first-level ciphertexts given as raw words
// ([count][2][k][n]) -> relinearize(multiply(a, b)) words ([count][2][k][n]), items split over `threads` workers.
int refshim_mul_relin_batch(void *ctx, const uint64_t *a, const uint64_t *b, void *rlk, uint64_t *out, uint64_t count,
                            int threads)
{
    try
    {
        auto &context = *reinterpret_cast<SEALContext *>(ctx);
        auto &keys = *reinterpret_cast<RelinKeys *>(rlk);
        auto cd = context.first_context_data();
        const size_t words = 2 * cd->parms().coeff_modulus().size() * cd->parms().poly_modulus_degree();
        Evaluator ev(context);
        std::vector<int> bad(threads, 0);
        std::vector<std::thread> ws;
        for (int t = 0; t < threads; t++)
            ws.emplace_back([&, t]() {
                try
                {
                    auto pool = MemoryPoolHandle::New();
                    Ciphertext ca(pool), cb(pool), d(pool);
                    ca.resize(context, context.first_parms_id(), 2);
                    cb.resize(context, context.first_parms_id(), 2);
                    for (uint64_t i = t; i < count; i += threads)
                    {
                        std::memcpy(ca.data(), a + i * words, words * sizeof(uint64_t));
                        std::memcpy(cb.data(), b + i * words, words * sizeof(uint64_t));
                        ev.multiply(ca, cb, d, pool);
                        ev.relinearize_inplace(d, keys, pool);
                        std::memcpy(out + i * words, d.data(), words * sizeof(uint64_t));
                    }
                }
                catch (...) { bad[t] = 1; }
            });
        for (auto &w : ws)
            w.join();
        for (int v : bad)
            if (v)
                return -1;
        return 0;
    }
    catch (...) { return -1; }
}
// forward transforms of `count` polynomials modulo `modulus`, split over `threads` workers (fully reduced output)
int refshim_ntt_forward_mt(uint64_t modulus, int logn, uint64_t *data, uint64_t count, int threads)
{
    try
    {
        auto *t = get_tables(modulus, logn);
        std::vector<std::thread> ws;
        for (int w = 0; w < threads; w++)
            ws.emplace_back([=]() {
                for (uint64_t i = w; i < count; i += threads)
                    ntt_negacyclic_harvey(CoeffIter(data + (i << logn)), *t);
            });
        for (auto &w : ws)
            w.join();
        return 0;
    }
    catch (...) { return -1; }
}

// --- CPU baseline: `threads` workers, each `iters` x (multiply + relinearize_inplace) on its own pool ------
// Returns wall seconds for the whole job (threads*iters operations) or <0 on error.
double refshim_time_mul_relin(void *ctx, void *a, void *b, void *rlk, int threads, int iters, int warmup)
{
    try
    {
        auto &context = *reinterpret_cast<SEALContext *>(ctx);
        auto &ca = *reinterpret_cast<Ciphertext *>(a);
        auto &cb = *reinterpret_cast<Ciphertext *>(b);
        auto &keys = *reinterpret_cast<RelinKeys *>(rlk);
        Evaluator ev(context);
        auto body = [&](int n) {
            auto pool = MemoryPoolHandle::New();
            Ciphertext d(pool);
            for (int i = 0; i < n; i++)
            {
                ev.multiply(ca, cb, d, pool);
                ev.relinearize_inplace(d, keys, pool);
            }
        };
        {
            std::vector<std::thread> ws;
            for (int t = 0; t < threads; t++)
                ws.emplace_back(body, warmup);
            for (auto &w : ws)
                w.join();
        }
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> ws;
        for (int t = 0; t < threads; t++)
            ws.emplace_back(body, iters);
        for (auto &w : ws)
            w.join();
        auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1 - t0).count();
    }
    catch (...) { return -1.0; }
}

// Same for batched forward+inverse NTT round trips of `count` polynomials, split across threads.
double refshim_time_ntt_roundtrip(uint64_t modulus, int logn, uint64_t *data, uint64_t count, int threads)
{
    try
    {
        auto *t = get_tables(modulus, logn);
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> ws;
        for (int w = 0; w < threads; w++)
            ws.emplace_back([=]() {
                for (uint64_t i = w; i < count; i += threads)
                {
                    ntt_negacyclic_harvey(CoeffIter(data + (i << logn)), *t);
                    inverse_ntt_negacyclic_harvey(CoeffIter(data + (i << logn)), *t);
                }
            });
        for (auto &w : ws)
            w.join();
        auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1 - t0).count();
    }
    catch (...) { return -1.0; }
}

} // extern "C"
