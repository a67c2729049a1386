/* oracle/bfv_oracle.h — TEST INFRASTRUCTURE ONLY (see bfv_oracle.c). */
#ifndef BFV_ORACLE_H
#define BFV_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
uint64_t orc_fnv1a64(const uint64_t *words, size_t count);
void orc_splitmix_fill(uint64_t *out, size_t count, uint64_t modulus, uint64_t *state);
#ifdef __cplusplus
}
#endif
#endif
