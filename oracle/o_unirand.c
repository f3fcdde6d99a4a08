/*
 * o_unirand.c -- oracle restatement of Polygon/unirand.zig.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference draws from std.crypto.random (unirand.zig:31), which cannot be
 * reproduced.  The draw *sequence* is kept (one draw for the offset, then one per
 * prime-table entry that passes the first two tests of the short-circuit `and` at
 * unirand.zig:42); the source of the draws is the documented splitmix64 stream of
 * include/myrenderer_b200.h.  PARITY UNPINNED for the random source itself.
 */
#include "mr_oracle.h"

/* unirand.zig:24 -- the table, in the reference's order (non-monotone tail included) */
static const uint32_t k_primes[] = {
    2,    3,    5,    7,    11,   13,   17,   19,   23,   29,   31,   37,   41,   43,
    47,   53,   59,   61,   67,   71,   73,   79,   83,   89,   97,   101,  103,  107,
    109,  113,  127,  131,  137,  139,  149,  151,  157,  163,  167,  173,  179,  181,
    191,  193,  197,  199,  211,  223,  227,  229,  233,  239,  241,  251,  257,  263,
    269,  271,  277,  281,  283,  293,  307,  311,  313,  317,  331,  337,  347,  349,
    353,  359,  367,  373,  379,  383,  389,  397,  401,  409,  419,  421,  431,  433,
    439,  443,  449,  457,  461,  463,  467,  479,  487,  491,  499,  503,  509,  521,
    523,  541,  601,  659,  733,  809,  863,  941,  1013, 1069, 1151, 1283, 1289, 1367,
    1447, 1499, 1579, 1637, 1723, 429494501u, 429493501u, 429486647u, 100001053u, 100002421u,
    10001567u};
#define K_NPRIMES (sizeof(k_primes) / sizeof(k_primes[0]))

uint64_t mr_o_rng_state0(uint64_t seed, uint64_t index) {
    return seed ^ (0x9E3779B97F4A7C15ull * (index + 1ull));
}

uint32_t mr_o_rng_u32(uint64_t* state) {
    uint64_t z;
    *state += 0x9E3779B97F4A7C15ull;
    z = *state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

/* unirand.zig:26-50 */
void mr_o_unirand_seed(mr_o_unirand* r, uint32_t top, uint64_t* rng_state) {
    uint32_t best = 1u;
    size_t i;
    r->at = 0; /* :29 */
    r->top = top; /* :33 */
    r->offset = 0; /* the reference leaves it undefined when top == 1 */
    if (top == 1u) { /* :34-37 */
        r->prime = 1u;
        return;
    }
    /* :38 -- u32 arithmetic; top == 0 wraps (top - 1) like ReleaseFast does */
    r->offset = mr_o_rng_u32(rng_state) % (uint32_t)(top - 1u) + 1u;
    for (i = 0; i < K_NPRIMES; ++i) { /* :41-45 */
        uint32_t p = k_primes[i];
        if (p < top && (top % p) != 0u) {
            if (mr_o_rng_u32(rng_state) % 3u > 0u) best = p; /* last passing entry wins */
        }
    }
    r->prime = best; /* :47 */
}

void mr_o_unirand_explicit(mr_o_unirand* r, uint32_t top, uint32_t offset, uint32_t prime) {
    r->at = 0;
    r->top = top;
    r->offset = offset;
    r->prime = prime;
}

/* unirand.zig:12-21 -- u32 arithmetic */
int mr_o_unirand_next(mr_o_unirand* r, uint32_t* out) {
    int have = 0;
    if (r->top > 0u && r->at < r->top) {
        *out = (uint32_t)(r->at * r->prime + r->offset) % r->top;
        have = 1;
    }
    r->at += 1u;
    return have;
}
