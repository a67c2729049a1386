/* oracle/bfv_oracle.h — TEST INFRASTRUCTURE ONLY (see bfv_oracle.c). */
#ifndef BFV_ORACLE_H
#define BFV_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXK 64

typedef struct orc_ctx
{
    size_t n;
    int logn;
    int K;                 /* key-level primes */
    uint64_t q[ORC_MAXK];  /* key-level primes (data level = first K-1, or K when K == 1) */
    uint64_t t;
    uint64_t m_sk, gamma, m_tilde;
    uint64_t aux[ORC_MAXK + 4]; /* get_primes(2n, 61, K+3): m_sk, gamma, B... */
} orc_ctx;

uint64_t orc_fnv1a64(const uint64_t *words, size_t count);
void orc_splitmix_fill(uint64_t *out, size_t count, uint64_t modulus, uint64_t *state);

int orc_ctx_init(orc_ctx *c, size_t n, const uint64_t *moduli, int count, uint64_t t);
int orc_is_prime(uint64_t v);
uint64_t orc_min_root(uint64_t p, size_t n); /* minimal primitive 2n-th root of unity */
/* aux base sizes for a data level with k residues: returns |B| (|Bsk| = |B|+1) */
int orc_base_b_size(const orc_ctx *c, int k);

/* transforms of one residue polynomial (canonical in / canonical out) */
void orc_ntt_forward(uint64_t *x, size_t n, uint64_t p);
void orc_ntt_inverse(uint64_t *x, size_t n, uint64_t p);

/* RNSTool steps with explicit bases (bsk = B primes then m_sk; rows n apart; m~ row last where present) */
void orc_fastbconv_m_tilde(const uint64_t *q, int k, const uint64_t *bsk, int nbsk, size_t n, const uint64_t *in, uint64_t *out);
void orc_sm_mrq(const uint64_t *q, int k, const uint64_t *bsk, int nbsk, size_t n, const uint64_t *in, uint64_t *out);
void orc_fast_floor(const uint64_t *q, int k, const uint64_t *bsk, int nbsk, size_t n, const uint64_t *in, uint64_t *out);
void orc_fastbconv_sk(const uint64_t *q, int k, const uint64_t *bsk, int nbsk, size_t n, const uint64_t *in, uint64_t *out);
uint64_t orc_aux_bases(size_t n, int nB, uint64_t *bsk_out);

/* ciphertext ops at a level with k residues; layouts as the reference: [poly][residue][coeff] */
void orc_add(const orc_ctx *c, int k, const uint64_t *a, const uint64_t *b, uint64_t *out, int size);
void orc_sub(const orc_ctx *c, int k, const uint64_t *a, const uint64_t *b, uint64_t *out, int size);
void orc_negate(const orc_ctx *c, int k, const uint64_t *a, uint64_t *out, int size);
int orc_multiply(const orc_ctx *c, int k, const uint64_t *a, int sa, const uint64_t *b, int sb, uint64_t *out);
/* key: [k digits][2][K][n] (NTT form); target: [k][n]; adds the switched pair into out (size-2, [2][k][n]) */
int orc_switch_key(const orc_ctx *c, int k, const uint64_t *target, const uint64_t *key, uint64_t *ct2);
int orc_relinearize(const orc_ctx *c, int k, const uint64_t *in3, const uint64_t *key, uint64_t *out2);
int orc_apply_galois(const orc_ctx *c, int k, const uint64_t *in2, uint32_t elt, const uint64_t *key, uint64_t *out2);
uint32_t orc_galois_elt_from_step(const orc_ctx *c, int steps);
int orc_multiply_plain(const orc_ctx *c, int k, const uint64_t *a, int size, const uint64_t *plain, size_t plain_count,
                       uint64_t *out);
int orc_add_plain(const orc_ctx *c, int k, const uint64_t *a, int size, const uint64_t *plain, size_t plain_count,
                  uint64_t *out, int subtract);
int orc_mod_switch_to_next(const orc_ctx *c, int k, const uint64_t *a, int size, uint64_t *out);
/* sk_ntt: secret key in NTT form, k residues [k][n]; plain_out: n coefficients mod t */
int orc_decrypt(const orc_ctx *c, int k, const uint64_t *ct, int size, const uint64_t *sk_ntt, uint64_t *plain_out);

#ifdef __cplusplus
}
#endif
#endif
