/*
 * Host-side helper for the PRIOR DRAWS of the lnZ_* functions: numpy's legacy global generator
 * (MT19937, `np.random.rand / randint / uniform`), continued in bulk from numpy's own state.
 *
 * The reference draws every prior sample from `np.random` (e.g. marginal_likelihoods.py:101-104)
 * and parity on identical host-drawn sample arrays requires the same stream.  numpy produces it
 * one 32-bit word at a time (~3.5 ns per word); the recurrence
 *     x[i+624] = x[i+397] ^ twist(x[i], x[i+1])
 * has a dependency distance of 227 words, so a long run of it vectorises.  The functions below
 * take the generator state (key[624], pos) as `np.random.get_state()` returns it, append the
 * raw state words of as many further blocks as the request needs, temper them into the same
 * output words numpy would hand out, and return the state for `np.random.set_state()`.
 * Bit-identical by construction (checked against numpy at first use and in the tests).
 * Build: gcc -O3 (libtriceratops_host.so).  Not a compute fallback: nothing of the light-curve
 * path lives here.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MT_N 624
#define MT_M 397
#define UPPER 0x80000000u
#define LOWER 0x7fffffffu

#if defined(__GNUC__) && defined(__x86_64__)
#define CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define CLONES
#endif

/* raw[0..624) holds the current key; fills raw[624 .. 624 + 624 * nblocks) with the state words
 * of the following blocks */
CLONES static void mt_extend(uint32_t* raw, int64_t nblocks) {
    const int64_t n = nblocks * MT_N;
    uint32_t* x = raw;
    /* chunks of at most 227 words keep every read behind the write front */
    for (int64_t base = 0; base < n; base += 224) {
        int64_t lim = base + 224 < n ? base + 224 : n;
        for (int64_t i = base; i < lim; i++) {
            uint32_t y = (x[i] & UPPER) | (x[i + 1] & LOWER);
            x[i + MT_N] = x[i + MT_M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
    }
}

static inline uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* numpy's mt19937_next_double over consecutive word pairs */
CLONES static void words_to_doubles(const uint32_t* w, double* out, int64_t n) {
    for (int64_t i = 0; i < n; i++) {
        uint32_t a = temper(w[2 * i]) >> 5, b = temper(w[2 * i + 1]) >> 6;
        out[i] = (a * 67108864.0 + b) / 9007199254740992.0;
    }
}

/* key[624], *pos (0..624): numpy's state, updated in place.  out[n] = np.random.rand(n). */
int trih_mt_rand(uint32_t* key, int32_t* pos, double* out, int64_t n) {
    if (n <= 0) return 0;
    const int64_t need = 2 * n;
    const int64_t have = MT_N - *pos;                       /* words left in the current key */
    const int64_t nblocks = need > have ? (need - have + MT_N - 1) / MT_N : 0;
    uint32_t* raw = (uint32_t*)malloc((size_t)(nblocks + 1) * MT_N * sizeof(uint32_t) + 64);
    if (!raw) return -1;
    memcpy(raw, key, MT_N * sizeof(uint32_t));
    mt_extend(raw, nblocks);
    if (out) words_to_doubles(raw + *pos, out, n);    /* (NULL: skip the draws) */
    int64_t cursor = *pos + need;                           /* in words from raw[0] */
    int64_t blk = cursor / MT_N, off = cursor % MT_N;
    if (off == 0 && blk > 0) { blk -= 1; off = MT_N; }      /* numpy regenerates lazily */
    memcpy(key, raw + blk * MT_N, MT_N * sizeof(uint32_t));
    *pos = (int32_t)off;
    free(raw);
    return 0;
}

/* out[n] = np.random.randint(low, low + rng + 1, n) for 0 < rng < 2^32 - 1 (legacy masked
 * rejection on single 32-bit words: numpy _bounded_integers, use_masked = True) */
int trih_mt_randint(uint32_t* key, int32_t* pos, int64_t low, uint32_t rng, int64_t* out,
                    int64_t n) {
    if (n <= 0) return 0;
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t* raw = (uint32_t*)malloc((size_t)2 * MT_N * sizeof(uint32_t));
    if (!raw) return -1;
    memcpy(raw, key, MT_N * sizeof(uint32_t));
    int64_t p = *pos;
    for (int64_t i = 0; i < n; i++) {
        uint32_t v;
        do {
            if (p == MT_N) {           /* next block, in place */
                mt_extend(raw, 1);
                memcpy(raw, raw + MT_N, MT_N * sizeof(uint32_t));
                p = 0;
            }
            v = temper(raw[p++]) & mask;
        } while (v > rng);
        out[i] = low + (int64_t)v;
    }
    memcpy(key, raw, MT_N * sizeof(uint32_t));
    *pos = (int32_t)p;
    free(raw);
    return 0;
}
