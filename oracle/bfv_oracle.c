/* oracle/bfv_oracle.c — TEST INFRASTRUCTURE ONLY: plain-C restatement of the reference's BFV hot path.
 * Nothing under sunscreen_b200/ links, loads or calls this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may.  (Restatement sections are appended below as they are pinned.) */
#include "bfv_oracle.h"

/* FNV-1a-64 over little-endian bytes (SURVEY.md App. B hash) */
uint64_t orc_fnv1a64(const uint64_t *words, size_t count)
{
    uint64_t h = 0xcbf29ce484222325ULL;
    for (size_t i = 0; i < count; i++)
    {
        uint64_t w = words[i];
        for (int b = 0; b < 8; b++)
        {
            h = (h ^ (w & 0xff)) * 0x100000001b3ULL;
            w >>= 8;
        }
    }
    return h;
}

/* splitmix64 stream of SURVEY.md App. B: every word = next() % modulus */
void orc_splitmix_fill(uint64_t *out, size_t count, uint64_t modulus, uint64_t *state)
{
    uint64_t s = *state;
    for (size_t i = 0; i < count; i++)
    {
        s += 0x9E3779B97F4A7C15ULL;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        out[i] = z % modulus;
    }
    *state = s;
}
