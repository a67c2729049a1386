// ksmac_tma.h — entry points of ksmac_tma.cu (key-switch inner product with a TMA-resident key tile).
#pragma once
// 1 when the tiled kernel can run this shape (and the driver exposes cuTensorMapEncodeTiled)
int b200_ksmac_tma_supported(long long n, int K);
// acc[item][c][I][n] = sum_J ks1[item][I][J][n] (*) key[J][c][key residue of I][n]; returns 0 or an error code
// (negative: tensor map could not be encoded; positive: cudaError_t of the launch)
int b200_ksmac_tma(int K, int fp, const void *primes /*PrimeDev[]*/, const void *fprimes /*NttPrimeFp[]*/, int special_idx, int key_rows,
                   const unsigned long long *ks1, const unsigned long long *key, unsigned long long *ks2, long long n, long long batch,
                   int sm_count, void *stream);
