// unirand.cuh -- device port of Polygon/unirand.zig (seed + next), shared with the host helper.
//
// unirand_seed (unirand.zig:26-50) draws from std.crypto.random; here the draws come from the
// documented counter-based stream of include/myrenderer_b200.h.  The draw *sequence* is the
// reference's: draw 0 feeds the offset (:38); every prime-table entry that passes
// `prime < top and top % prime != 0` consumes the next draw (short-circuit `and`, :42); the last
// entry whose draw % 3 > 0 wins (:43).  Because the stream is counter-based, draw k is
// mix(state0 + (k+1)*GOLDEN) and the table walk parallelises over lanes: a ballot gives each
// passing entry its draw index, another ballot finds the last winner.
#pragma once
#include <cstdint>

#define MR_GOLDEN 0x9E3779B97F4A7C15ull
#define MR_NPRIMES 123

// unirand.zig:24, in the reference's order -- the one product copy of the table (device array here, host array in api.cu)
#define MR_PRIME_LIST                                                                                       \
    2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 101, 103, 107,  \
    109, 113, 127, 131, 137, 139, 149, 151, 157, 163, 167, 173, 179, 181, 191, 193, 197, 199, 211, 223, 227, 229,   \
    233, 239, 241, 251, 257, 263, 269, 271, 277, 281, 283, 293, 307, 311, 313, 317, 331, 337, 347, 349, 353, 359,   \
    367, 373, 379, 383, 389, 397, 401, 409, 419, 421, 431, 433, 439, 443, 449, 457, 461, 463, 467, 479, 487, 491,   \
    499, 503, 509, 521, 523, 541, 601, 659, 733, 809, 863, 941, 1013, 1069, 1151, 1283, 1289, 1367, 1447, 1499,     \
    1579, 1637, 1723, 429494501u, 429493501u, 429486647u, 100001053u, 100002421u, 10001567u
__device__ __constant__ uint32_t k_unirand_primes_dev[MR_NPRIMES] = {MR_PRIME_LIST};

__host__ __device__ __forceinline__ uint64_t mr_rng_state0_hd(uint64_t seed, uint64_t index) {
    return seed ^ (MR_GOLDEN * (index + 1ull));
}

// draw number k (0-based) of the stream that starts at state0
__host__ __device__ __forceinline__ uint32_t mr_rng_draw(uint64_t state0, uint32_t k) {
    uint64_t z = state0 + MR_GOLDEN * (uint64_t)(k + 1u);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

// All 32 lanes call this with the same arguments; every lane returns the same (offset, prime).
__device__ __forceinline__ void unirand_seed_warp(uint32_t top, uint64_t seed, uint64_t index, uint32_t lane,
                                                  uint32_t* offset_out, uint32_t* prime_out) {
    if (top == 1u) {  // unirand.zig:34-37 (offset is left undefined there; 0 here)
        *offset_out = 0u;
        *prime_out = 1u;
        return;
    }
    const uint64_t s0 = mr_rng_state0_hd(seed, index);
    *offset_out = mr_rng_draw(s0, 0u) % (uint32_t)(top - 1u) + 1u;  // :38, u32 wrap for top == 0
    uint32_t draws = 1u;  // draws consumed so far
    uint32_t best = 1u;
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t b = 0; b < MR_NPRIMES; b += 32) {
        const uint32_t i = b + lane;
        const uint32_t p = i < MR_NPRIMES ? k_unirand_primes_dev[i] : 0u;
        const bool pass = (i < MR_NPRIMES) && (p < top) && (top % p != 0u);
        const uint32_t pm = __ballot_sync(0xFFFFFFFFu, pass);
        bool win = false;
        if (pass) win = (mr_rng_draw(s0, draws + __popc(pm & lt)) % 3u) > 0u;
        const uint32_t wm = __ballot_sync(0xFFFFFFFFu, win);
        if (wm) best = __shfl_sync(0xFFFFFFFFu, p, 31 - __clz(wm));  // last winner in table order
        draws += __popc(pm);
    }
    *prime_out = best;
}
