// sampling.cpp — see sampling.h.
#include "sampling.h"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <random>
#include <stdexcept>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define B200_X86_SIMD 1
#endif

namespace b200
{
namespace
{
typedef unsigned __int128 u128;
const uint64_t IV[8] = { 0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                         0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL };
const unsigned char SIGMA[12][16] = {
    { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 },
    { 11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4 }, { 7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8 },
    { 9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13 }, { 2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9 },
    { 12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11 }, { 13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10 },
    { 6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5 }, { 10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0 },
    { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 }
};
inline uint64_t rotr(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }

struct B2State
{
    uint64_t h[8];
    u128 t = 0;
    unsigned char buf[128];
    size_t buflen = 0;
    size_t outlen = 64;
};

void compress(B2State &S, const unsigned char *block, bool last)
{
    uint64_t m[16], v[16];
    std::memcpy(m, block, 128);
    for (int i = 0; i < 8; i++)
    {
        v[i] = S.h[i];
        v[i + 8] = IV[i];
    }
    v[12] ^= (uint64_t)S.t;
    v[13] ^= (uint64_t)(S.t >> 64);
    if (last)
        v[14] = ~v[14];
    for (int r = 0; r < 12; r++)
    {
        const unsigned char *s = SIGMA[r];
        auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
            v[a] = v[a] + v[b] + x;
            v[d] = rotr(v[d] ^ v[a], 32);
            v[c] = v[c] + v[d];
            v[b] = rotr(v[b] ^ v[c], 24);
            v[a] = v[a] + v[b] + y;
            v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];
            v[b] = rotr(v[b] ^ v[c], 63);
        };
        G(0, 4, 8, 12, m[s[0]], m[s[1]]);
        G(1, 5, 9, 13, m[s[2]], m[s[3]]);
        G(2, 6, 10, 14, m[s[4]], m[s[5]]);
        G(3, 7, 11, 15, m[s[6]], m[s[7]]);
        G(0, 5, 10, 15, m[s[8]], m[s[9]]);
        G(1, 6, 11, 12, m[s[10]], m[s[11]]);
        G(2, 7, 8, 13, m[s[12]], m[s[13]]);
        G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++)
        S.h[i] ^= v[i] ^ v[i + 8];
}

// 64-byte BLAKE2b parameter block
void init_param(B2State &S, const unsigned char P[64])
{
    for (int i = 0; i < 8; i++)
    {
        uint64_t w;
        std::memcpy(&w, P + 8 * i, 8);
        S.h[i] = IV[i] ^ w;
    }
    S.t = 0;
    S.buflen = 0;
    S.outlen = P[0];
}
void update(B2State &S, const unsigned char *in, size_t len)
{
    while (len)
    {
        if (S.buflen == 128)
        { // buffer full and more input follows: compress it (never the last block here)
            S.t += 128;
            compress(S, S.buf, false);
            S.buflen = 0;
        }
        size_t take = std::min(len, (size_t)128 - S.buflen);
        std::memcpy(S.buf + S.buflen, in, take);
        S.buflen += take;
        in += take;
        len -= take;
    }
}
void final(B2State &S, unsigned char *out, size_t outlen)
{
    S.t += S.buflen;
    std::memset(S.buf + S.buflen, 0, 128 - S.buflen);
    compress(S, S.buf, true);
    unsigned char full[64];
    std::memcpy(full, S.h, 64);
    std::memcpy(out, full, outlen);
}
void store32(unsigned char *p, uint32_t v) { std::memcpy(p, &v, 4); }

#ifdef B200_X86_SIMD
// BLAKE2Xb expansion nodes are independent BLAKE2b compressions of the same 64-byte root that differ only in the node
// offset of their parameter block, so W of them run side by side in the 64-bit lanes of one vector register (the
// generator behind every key and every encryption is otherwise the host-side bottleneck: one scalar compression per 64
// output bytes).  h1_xor = the parameter word that carries the node offset, for lane 0; lane j adds j to the offset.
#define B200_G(a, b, c, d, x, y)                                                                                       \
    a = ADD(ADD(a, b), x);                                                                                             \
    d = ROR(XOR(d, a), 32);                                                                                            \
    c = ADD(c, d);                                                                                                     \
    b = ROR(XOR(b, c), 24);                                                                                            \
    a = ADD(ADD(a, b), y);                                                                                             \
    d = ROR(XOR(d, a), 16);                                                                                            \
    c = ADD(c, d);                                                                                                     \
    b = ROR(XOR(b, c), 63);
#define B200_ROUNDS()                                                                                                  \
    for (int r = 0; r < 12; r++)                                                                                       \
    {                                                                                                                  \
        const unsigned char *s = SIGMA[r];                                                                             \
        B200_G(v0, v4, v8, v12, m[s[0]], m[s[1]])                                                                       \
        B200_G(v1, v5, v9, v13, m[s[2]], m[s[3]])                                                                       \
        B200_G(v2, v6, v10, v14, m[s[4]], m[s[5]])                                                                      \
        B200_G(v3, v7, v11, v15, m[s[6]], m[s[7]])                                                                      \
        B200_G(v0, v5, v10, v15, m[s[8]], m[s[9]])                                                                      \
        B200_G(v1, v6, v11, v12, m[s[10]], m[s[11]])                                                                    \
        B200_G(v2, v7, v8, v13, m[s[12]], m[s[13]])                                                                     \
        B200_G(v3, v4, v9, v14, m[s[14]], m[s[15]])                                                                     \
    }

__attribute__((target("avx512f"))) void expand_nodes_avx512(const uint64_t h0[8], const uint64_t root[8], uint64_t node0,
                                                            unsigned char *out)
{
    typedef __m512i V;
#define ADD _mm512_add_epi64
#define XOR _mm512_xor_si512
#define ROR(x, n) _mm512_ror_epi64(x, n)
    V m[16];
    for (int i = 0; i < 8; i++)
        m[i] = _mm512_set1_epi64((long long)root[i]);
    for (int i = 8; i < 16; i++)
        m[i] = _mm512_setzero_si512();
    // node offset: low 32 bits of parameter word 1 (bytes 8..11)
    const V lane = _mm512_set_epi64(7, 6, 5, 4, 3, 2, 1, 0);
    V h[8];
    for (int i = 0; i < 8; i++)
        h[i] = _mm512_set1_epi64((long long)h0[i]);
    h[1] = XOR(h[1], ADD(_mm512_set1_epi64((long long)node0), lane));
    V v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    V v8 = _mm512_set1_epi64((long long)IV[0]), v9 = _mm512_set1_epi64((long long)IV[1]), v10 = _mm512_set1_epi64((long long)IV[2]),
      v11 = _mm512_set1_epi64((long long)IV[3]), v12 = _mm512_set1_epi64((long long)(IV[4] ^ 64)),
      v13 = _mm512_set1_epi64((long long)IV[5]), v14 = _mm512_set1_epi64((long long)~IV[6]), v15 = _mm512_set1_epi64((long long)IV[7]);
    B200_ROUNDS()
    alignas(64) uint64_t t[8][8];
    _mm512_store_si512((void *)t[0], XOR(h[0], XOR(v0, v8)));
    _mm512_store_si512((void *)t[1], XOR(h[1], XOR(v1, v9)));
    _mm512_store_si512((void *)t[2], XOR(h[2], XOR(v2, v10)));
    _mm512_store_si512((void *)t[3], XOR(h[3], XOR(v3, v11)));
    _mm512_store_si512((void *)t[4], XOR(h[4], XOR(v4, v12)));
    _mm512_store_si512((void *)t[5], XOR(h[5], XOR(v5, v13)));
    _mm512_store_si512((void *)t[6], XOR(h[6], XOR(v6, v14)));
    _mm512_store_si512((void *)t[7], XOR(h[7], XOR(v7, v15)));
    for (int j = 0; j < 8; j++)
    {
        uint64_t w[8];
        for (int i = 0; i < 8; i++)
            w[i] = t[i][j];
        std::memcpy(out + j * 64, w, 64);
    }
#undef ADD
#undef XOR
#undef ROR
}

__attribute__((target("avx2"))) void expand_nodes_avx2(const uint64_t h0[8], const uint64_t root[8], uint64_t node0, unsigned char *out)
{
    typedef __m256i V;
#define ADD _mm256_add_epi64
#define XOR _mm256_xor_si256
#define ROR(x, n) _mm256_or_si256(_mm256_srli_epi64(x, n), _mm256_slli_epi64(x, 64 - (n)))
    V m[16];
    for (int i = 0; i < 8; i++)
        m[i] = _mm256_set1_epi64x((long long)root[i]);
    for (int i = 8; i < 16; i++)
        m[i] = _mm256_setzero_si256();
    const V lane = _mm256_set_epi64x(3, 2, 1, 0);
    V h[8];
    for (int i = 0; i < 8; i++)
        h[i] = _mm256_set1_epi64x((long long)h0[i]);
    h[1] = XOR(h[1], ADD(_mm256_set1_epi64x((long long)node0), lane));
    V v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    V v8 = _mm256_set1_epi64x((long long)IV[0]), v9 = _mm256_set1_epi64x((long long)IV[1]), v10 = _mm256_set1_epi64x((long long)IV[2]),
      v11 = _mm256_set1_epi64x((long long)IV[3]), v12 = _mm256_set1_epi64x((long long)(IV[4] ^ 64)),
      v13 = _mm256_set1_epi64x((long long)IV[5]), v14 = _mm256_set1_epi64x((long long)~IV[6]), v15 = _mm256_set1_epi64x((long long)IV[7]);
    B200_ROUNDS()
    alignas(32) uint64_t t[8][4];
    _mm256_store_si256((__m256i *)t[0], XOR(h[0], XOR(v0, v8)));
    _mm256_store_si256((__m256i *)t[1], XOR(h[1], XOR(v1, v9)));
    _mm256_store_si256((__m256i *)t[2], XOR(h[2], XOR(v2, v10)));
    _mm256_store_si256((__m256i *)t[3], XOR(h[3], XOR(v3, v11)));
    _mm256_store_si256((__m256i *)t[4], XOR(h[4], XOR(v4, v12)));
    _mm256_store_si256((__m256i *)t[5], XOR(h[5], XOR(v5, v13)));
    _mm256_store_si256((__m256i *)t[6], XOR(h[6], XOR(v6, v14)));
    _mm256_store_si256((__m256i *)t[7], XOR(h[7], XOR(v7, v15)));
    for (int j = 0; j < 4; j++)
    {
        uint64_t w[8];
        for (int i = 0; i < 8; i++)
            w[i] = t[i][j];
        std::memcpy(out + j * 64, w, 64);
    }
#undef ADD
#undef XOR
#undef ROR
}
int simd_width()
{
    static const int w = [] {
        __builtin_cpu_init();
        const int hw = __builtin_cpu_supports("avx512f") ? 8 : __builtin_cpu_supports("avx2") ? 4 : 1;
        const char *e = std::getenv("B200_PRNG_SIMD"); // developer/test knob: cap the width (1 = scalar, 4 = AVX2)
        const int cap = e ? std::atoi(e) : 8;
        return hw >= 8 && cap >= 8 ? 8 : hw >= 4 && cap >= 4 ? 4 : 1;
    }();
    return w;
}
#endif
} // namespace

void blake2xb(void *out_, size_t outlen, const void *in, size_t inlen, const void *key, size_t keylen)
{
    unsigned char *out = (unsigned char *)out_;
    // root hash H0: BLAKE2b-512 with xof_length = outlen
    unsigned char P[64];
    std::memset(P, 0, 64);
    P[0] = 64;
    P[1] = (unsigned char)keylen;
    P[2] = 1;
    P[3] = 1;
    store32(P + 12, (uint32_t)outlen);
    B2State S;
    init_param(S, P);
    if (keylen)
    {
        unsigned char block[128];
        std::memset(block, 0, 128);
        std::memcpy(block, key, keylen);
        update(S, block, 128);
    }
    update(S, (const unsigned char *)in, inlen);
    unsigned char root[64];
    final(S, root, 64);
    // expansion nodes
    std::memset(P, 0, 64);
    P[1] = 0;
    P[2] = 0;
    P[3] = 0;
    store32(P + 4, 64); // leaf_length
    store32(P + 12, (uint32_t)outlen);
    P[16] = 0;  // node_depth
    P[17] = 64; // inner_length
    size_t i = 0;
#ifdef B200_X86_SIMD
    if (const int W = simd_width(); W > 1)
    { // full 64-byte nodes, W at a time
        P[0] = 64;
        store32(P + 8, 0);
        uint64_t h0[8], rootw[8];
        for (int j = 0; j < 8; j++)
        {
            uint64_t w;
            std::memcpy(&w, P + 8 * j, 8);
            h0[j] = IV[j] ^ w;
        }
        std::memcpy(rootw, root, 64);
        for (; outlen >= (size_t)W * 64; i += W, outlen -= (size_t)W * 64)
        {
            if (W == 8)
                expand_nodes_avx512(h0, rootw, i, out + i * 64);
            else
                expand_nodes_avx2(h0, rootw, i, out + i * 64);
        }
    }
#endif
    for (; outlen > 0; i++)
    {
        size_t block = outlen < 64 ? outlen : 64;
        P[0] = (unsigned char)block;
        store32(P + 8, (uint32_t)i); // node_offset
        B2State C;
        init_param(C, P);
        update(C, root, 64);
        final(C, out + i * 64, block);
        outlen -= block;
    }
}

Blake2xbPrng::Blake2xbPrng(const PrngSeed &seed) : seed_(seed), buffer_(4096), head_(4096) {}

void Blake2xbPrng::refill()
{
    blake2xb(buffer_.data(), buffer_.size(), &counter_, sizeof(counter_), seed_.data(), seed_.size() * sizeof(uint64_t));
    counter_++;
}

void Blake2xbPrng::generate(size_t bytes, void *dst_)
{
    unsigned char *dst = (unsigned char *)dst_;
    // the buffer starts "empty" (head at end) and is refilled whenever the head reaches the end — after the copy
    // (S/randomgen.cpp:176-193); the first call therefore refills before copying anything.
    while (bytes)
    {
        if (head_ == buffer_.size())
        {
            refill();
            head_ = 0;
        }
        size_t cur = std::min(bytes, buffer_.size() - head_);
        std::memcpy(dst, buffer_.data() + head_, cur);
        head_ += cur;
        dst += cur;
        bytes -= cur;
    }
}

PrngSeed random_seed()
{
    std::random_device rd;
    PrngSeed s;
    for (auto &w : s)
        w = ((uint64_t)rd() << 32) | rd();
    return s;
}

void sample_poly_ternary(Blake2xbPrng &prng, size_t n, const std::vector<uint64_t> &moduli, uint64_t *out)
{
    PrngEngine engine(prng);
    std::uniform_int_distribution<uint64_t> dist(0, 2);
    for (size_t c = 0; c < n; c++)
    {
        uint64_t r = dist(engine);
        uint64_t flag = (uint64_t)(-(int64_t)(r == 0));
        for (size_t i = 0; i < moduli.size(); i++)
            out[i * n + c] = r + (flag & moduli[i]) - 1;
    }
}

void sample_poly_normal(Blake2xbPrng &prng, size_t n, const std::vector<uint64_t> &moduli, uint64_t *out)
{
    PrngEngine engine(prng);
    std::normal_distribution<double> normal(0.0, 3.2); // noise_standard_deviation (S/util/globals.h:36)
    const double max_dev = 3.2 * 6;                     // noise_max_deviation
    for (size_t c = 0; c < n; c++)
    {
        double value;
        do
        {
            value = normal(engine);
        } while (std::abs(value - 0.0) > max_dev);
        int64_t noise = (int64_t)value;
        uint64_t flag = (uint64_t)(-(int64_t)(noise < 0));
        for (size_t i = 0; i < moduli.size(); i++)
            out[i * n + c] = (uint64_t)noise + (flag & moduli[i]);
    }
}

void sample_poly_uniform(Blake2xbPrng &prng, size_t n, const std::vector<uint64_t> &moduli, uint64_t *out)
{
    prng.generate(moduli.size() * n * sizeof(uint64_t), out);
    const uint64_t max_random = 0xFFFFFFFFFFFFFFFFULL;
    for (size_t j = 0; j < moduli.size(); j++)
    {
        const uint64_t q = moduli[j];
        const uint64_t max_multiple = max_random - (max_random % q) - 1;
        const uint64_t qinv = (uint64_t)(((u128)1 << 64) / q); // floor(2^64 / q): r % q without a division per word
        for (size_t c = 0; c < n; c++)
        {
            uint64_t r = out[j * n + c];
            while (r >= max_multiple)
                prng.generate(sizeof(r), &r);
            uint64_t rem = r - (uint64_t)(((u128)r * qinv) >> 64) * q; // quotient estimate is at most 1 too small
            while (rem >= q)
                rem -= q;
            out[j * n + c] = rem;
        }
    }
}
} // namespace b200
