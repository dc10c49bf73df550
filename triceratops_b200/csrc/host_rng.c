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
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MT_N 624
#define MT_M 397
#define UPPER 0x80000000u
#define LOWER 0x7fffffffu

#if defined(__GNUC__) && defined(__x86_64__)
#define CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#define CPU_RELAX() __builtin_ia32_pause()
#else
#define CLONES
#define CPU_RELAX() ((void)0)
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

/* Grow-only scratch for the raw state words, one per calling thread: a fresh 8-40 MB block per
 * call would be page-faulted in again every time (a third of a bulk rand(1e6)). */
static __thread uint32_t* t_raw = NULL;
static __thread size_t t_raw_cap = 0;

static uint32_t* raw_scratch(size_t words) {
    if (words > t_raw_cap) {
        free(t_raw);
        t_raw_cap = words + words / 8;
        t_raw = (uint32_t*)malloc(t_raw_cap * sizeof(uint32_t));
        if (!t_raw) t_raw_cap = 0;
    }
    return t_raw;
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

/* key[624], *pos (0..624): numpy's state, updated in place.  out[n] = np.random.rand(n).
 * The state recurrence is sequential; the tempering + conversion of the words is spread over
 * `nthreads` (the caller's choice per call; large n only). */
int trih_mt_rand(uint32_t* key, int32_t* pos, double* out, int64_t n, int nthreads) {
    if (n <= 0) return 0;
    const int64_t need = 2 * n;
    const int64_t have = MT_N - *pos;                       /* words left in the current key */
    const int64_t nblocks = need > have ? (need - have + MT_N - 1) / MT_N : 0;
    uint32_t* raw = raw_scratch((size_t)(nblocks + 1) * MT_N + 16);
    if (!raw) return -1;
    memcpy(raw, key, MT_N * sizeof(uint32_t));
    mt_extend(raw, nblocks);
    if (out) {                                        /* (NULL: skip the draws) */
        const uint32_t* w = raw + *pos;
        if (nthreads > 1 && n >= (1 << 18)) {
            const int64_t per = (n + nthreads - 1) / nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads)
            for (int t = 0; t < nthreads; t++) {
                const int64_t lo = t * per, hi = lo + per < n ? lo + per : n;
                if (lo < hi) words_to_doubles(w + 2 * lo, out + lo, hi - lo);
            }
        } else {
            words_to_doubles(w, out, n);
        }
    }
    int64_t cursor = *pos + need;                           /* in words from raw[0] */
    int64_t blk = cursor / MT_N, off = cursor % MT_N;
    if (off == 0 && blk > 0) { blk -= 1; off = MT_N; }      /* numpy regenerates lazily */
    memcpy(key, raw + blk * MT_N, MT_N * sizeof(uint32_t));
    *pos = (int32_t)off;
    return 0;
}

/* out[n] = np.random.randint(low, low + rng + 1, n) for 0 < rng < 2^32 - 1 (legacy masked
 * rejection on single 32-bit words: numpy _bounded_integers, use_masked = True).  Every word of
 * the stream is either accepted or rejected on its own, so the words are produced in bulk and
 * compacted branch-free: the output is written for every word and the cursor only advances past
 * the accepted ones (numpy's own loop costs ~10 ns per word in the block-boundary test). */
static int64_t g_randint_margin = 64;
void trih_debug_randint_margin(int64_t words) { g_randint_margin = words; }   /* (tests) */

static int64_t compact_accepted(const uint32_t* raw, int64_t* cursor, int64_t end, uint32_t mask,
                                uint32_t rng, int64_t low, int64_t* out, int64_t k, int64_t n) {
    int64_t i = *cursor;
    /* out[n - 1] is the last slot that may be written: stop as soon as k reaches n */
    for (; i < end && k < n; i++) {
        const uint32_t v = temper(raw[i]) & mask;
        out[k] = low + (int64_t)v;
        k += (v <= rng);
    }
    *cursor = i;
    return k;
}

int trih_mt_randint(uint32_t* key, int32_t* pos, int64_t low, uint32_t rng, int64_t* out,
                    int64_t n) {
    if (n <= 0) return 0;
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    const double p = ((double)rng + 1.0) / ((double)mask + 1.0);       /* acceptance, >= 1/2 */
    int64_t est = (int64_t)((double)n / p + 8.0 * sqrt((double)n * (1.0 - p)) / p)
                  + g_randint_margin;
    if (est < 1) est = 1;
    const int64_t have = MT_N - *pos;
    int64_t nblocks = est > have ? (est - have + MT_N - 1) / MT_N : 0;
    const int64_t more = 8;                                  /* blocks per refill */
    uint32_t* raw = raw_scratch((size_t)((nblocks > more ? nblocks : more) + 2) * MT_N + 16);
    if (!raw) return -1;
    memcpy(raw, key, MT_N * sizeof(uint32_t));
    mt_extend(raw, nblocks);
    int64_t cursor = *pos, end = (nblocks + 1) * MT_N, k = 0;
    for (;;) {
        k = compact_accepted(raw, &cursor, end, mask, rng, low, out, k, n);
        if (k == n) break;
        /* the estimate fell short: the last block becomes the key, a few more follow */
        memmove(raw, raw + end - MT_N, MT_N * sizeof(uint32_t));
        mt_extend(raw, more);
        cursor = MT_N;
        end = (more + 1) * MT_N;
    }
    int64_t blk = cursor / MT_N, off = cursor % MT_N;
    if (off == 0 && blk > 0) { blk -= 1; off = MT_N; }      /* numpy regenerates lazily */
    memcpy(key, raw + blk * MT_N, MT_N * sizeof(uint32_t));
    *pos = (int32_t)off;
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * scipy.stats.beta.rvs(a, b, size=n) with a < 1 < b on numpy's legacy generator: the
 * eccentricity prior of planets, Beta(0.867, 3.03) (reference priors.py:148).  numpy's
 * legacy_beta draws Ga ~ Gamma(a) (rejection, two uniforms per attempt) and Gb ~ Gamma(b)
 * (Marsaglia-Tsang on polar-method gaussians, one of every pair cached in the generator state)
 * and returns Ga / (Ga + Gb): ~110 ns per sample of log / pow / sqrt, and sequential, because
 * every accept / reject decides where in the stream the next sample starts.  It is the largest
 * single item of a calc_probs call's host time (6 x 1e6 samples).
 *
 * The state of the sampler at a sample boundary is small: (position in the stream, which polar
 * pair -- if any -- left a gaussian in the cache).  Two walks that reach the SAME state produce
 * the same samples from there on.  So the stream is cut into chunks by position; every chunk is
 * walked by its own thread with numpy's exact algorithm, starting from the guess "a sample
 * starts here, cache empty"; afterwards the chunks are stitched in order: the true state at the
 * start of a chunk is walked forward until it coincides with a boundary state the chunk's
 * thread recorded (a few dozen samples: both walks hit every ~4.4th double, and the cache
 * empties every other gaussian), and the thread's samples are adopted from there.  Same
 * operations, same libm calls, no contraction: outputs and final generator state (position,
 * cached gaussian) are bit-identical to numpy's; tests/test_fastrng.py and the self-check of
 * _fastrng.py compare them.
 * ------------------------------------------------------------------------------------------- */
#include <math.h>
#include <stdio.h>
#include <time.h>

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

static inline double dbl_at(const uint32_t* w, int64_t m) {
    uint32_t a = temper(w[2 * m]) >> 5, b = temper(w[2 * m + 1]) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

typedef struct {
    const uint32_t* w;    /* raw state words from the current position */
    int64_t cap;          /* doubles available */
    double a, bb, c;      /* shape a; b - 1/3; 1 / sqrt(9 (b - 1/3)) */
} beta_ctx;

#define SRC_NONE (-2)     /* no cached gaussian */
#define SRC_ENTRY (-1)    /* the gaussian cached in the generator state on entry */

typedef struct {
    int64_t m;            /* next double of the stream */
    int64_t src;          /* SRC_NONE, SRC_ENTRY, or the first double of the pair that cached it */
    double cache;         /* the cached gaussian (valid when src != SRC_NONE) */
} walk_state;

/* numpy legacy_gauss */
static inline int walk_gauss(const beta_ctx* C, walk_state* S, double* g) {
    if (S->src != SRC_NONE) {
        *g = S->cache;
        S->src = SRC_NONE;
        S->cache = 0.0;
        return 0;
    }
    double f, x1, x2, r2;
    int64_t q;
    do {
        if (S->m + 2 > C->cap) return -2;
        q = S->m;
        x1 = 2.0 * dbl_at(C->w, q) - 1.0;
        x2 = 2.0 * dbl_at(C->w, q + 1) - 1.0;
        r2 = x1 * x1 + x2 * x2;
        S->m += 2;
    } while (r2 >= 1.0 || r2 == 0.0);
    f = sqrt(-2.0 * log(r2) / r2);
    S->cache = f * x1;
    S->src = q;
    *g = f * x2;
    return 0;
}

/* numpy legacy_beta for a < 1 < b: one sample from state S */
static inline int walk_beta(const beta_ctx* C, walk_state* S, double* val) {
    const double shape = C->a, b = C->bb, c = C->c;
    double U, V, X, Y, Ga, Gb;
    for (;;) {                                   /* legacy_standard_gamma, shape < 1 */
        if (S->m + 2 > C->cap) return -2;
        U = dbl_at(C->w, S->m);
        V = -log(1.0 - dbl_at(C->w, S->m + 1));
        S->m += 2;
        if (U <= 1.0 - shape) {
            X = pow(U, 1. / shape);
            if (X <= V) break;
        } else {
            Y = -log((1 - U) / shape);
            X = pow(1.0 - shape + shape * Y, 1. / shape);
            if (X <= (V + Y)) break;
        }
    }
    Ga = X;
    for (;;) {                                   /* legacy_standard_gamma, shape > 1 */
        do {
            if (walk_gauss(C, S, &X)) return -2;
            V = 1.0 + c * X;
        } while (V <= 0.0);
        V = V * V * V;
        if (S->m + 1 > C->cap) return -2;
        U = dbl_at(C->w, S->m);
        S->m += 1;
        if (U < 1.0 - 0.0331 * (X * X) * (X * X)) break;
        if (log(U) < 0.5 * X * X + b * (1. - V + log(V))) break;
    }
    Gb = b * V;
    *val = Ga / (Ga + Gb);
    return 0;
}

typedef struct {
    int64_t n;            /* samples recorded */
    int64_t cap;
    int64_t* m0;          /* state at the start of sample k */
    int64_t* src0;
    double* val;
    walk_state end;       /* state after the last sample */
    int rc;
} chunk_rec;

int trih_legacy_beta(uint32_t* key, int32_t* pos, int32_t* has_gauss, double* gauss, double a,
                     double b, double* out, int64_t n, int nthreads) {
    if (!(a > 0.0 && a < 1.0 && b > 1.0) || n < 0) return -3;
    if (n == 0) return 0;
    if (nthreads < 1) nthreads = 1;
    const int trace = getenv("TRI_B200_RNG_TRACE") != NULL;
    const double t_start = now_s();
    double t_pilot = 0, t_extend = 0, t_walk = 0;
    beta_ctx C;
    C.a = a;
    C.bb = b - 1. / 3.;
    C.c = 1. / sqrt(9 * C.bb);
    /* doubles per sample from a pilot walk over the first block(s) of the stream */
    double per = 4.5;
    {
        uint32_t pilot[9 * MT_N];
        memcpy(pilot, key, MT_N * sizeof(uint32_t));
        mt_extend(pilot, 8);
        C.w = pilot + *pos;
        C.cap = (9 * MT_N - *pos) / 2;
        walk_state S = {0, SRC_NONE, 0.0};
        int64_t k = 0;
        double v;
        while (k < n && walk_beta(&C, &S, &v) == 0) k++;
        if (k >= 64) per = (double)S.m / (double)k;
    }
    t_pilot = now_s();
    /* the stream: estimated need + 2 % + slack for the sequential tail */
    const int64_t est = (int64_t)(per * (double)n * 1.02) + 4096;
    const int64_t cap = est + 65536;
    const int64_t need = 2 * cap;
    const int64_t have = MT_N - *pos;
    const int64_t nblocks = need > have ? (need - have + MT_N - 1) / MT_N : 0;
    uint32_t* raw = raw_scratch((size_t)(nblocks + 1) * MT_N + 16);
    if (!raw) return -1;
    memcpy(raw, key, MT_N * sizeof(uint32_t));
    /* only the head of the stream is produced up front; thread 0 produces the rest while the
     * other threads already walk the first chunks (see the parallel region) */
    const int64_t piece = 1024;                              /* blocks per publication */
    int64_t ext_blocks = nblocks < 2 * piece || nthreads < 2 ? nblocks : 2 * piece;
    mt_extend(raw, ext_blocks);
    C.w = raw + *pos;
    C.cap = cap;
    t_extend = now_s();

    int nchunks = (int)(n / 16384);
    if (nchunks > 8 * nthreads) nchunks = 8 * nthreads;
    if (nchunks < 1) nchunks = 1;
    chunk_rec* R = (chunk_rec*)calloc((size_t)nchunks, sizeof(chunk_rec));
    if (!R) return -1;
    const walk_state entry = {0, *has_gauss ? SRC_ENTRY : SRC_NONE, *has_gauss ? *gauss : 0.0};
    int rc = 0;

    /* (position + [a gaussian is cached]) mod 2 never changes along a walk: a sample consumes
     * 2 doubles per gamma(a) attempt, 2 per polar attempt and 1 per Marsaglia-Tsang attempt,
     * and every Marsaglia-Tsang attempt also toggles the cache.  A guess of the other class
     * could never meet the true walk, so the chunks start on the true walk's class.  (The one
     * exception -- a gaussian below -1/c = -4.9 is redrawn without a uniform, 4e-7 per sample --
     * flips the class: the stitcher notices and the remaining chunks are walked again.) */
    int cls = (int)((entry.m + (entry.src != SRC_NONE)) & 1);
    int first = 0, t = 0;
    walk_state S = entry;             /* the true walk (stitcher) */
    int64_t count = 0;
    int next_chunk;
walk_again:
    next_chunk = first;
#pragma omp parallel num_threads(nthreads)
    {
    if (omp_get_thread_num() == 0) {
        /* the producer: the rest of the stream, published piece by piece */
        for (int64_t b = __atomic_load_n(&ext_blocks, __ATOMIC_RELAXED); b < nblocks;) {
            const int64_t k = nblocks - b < piece ? nblocks - b : piece;
            mt_extend(raw + b * MT_N, k);
            b += k;
            __atomic_store_n(&ext_blocks, b, __ATOMIC_RELEASE);
        }
    }
    for (;;) {
        const int t = __atomic_fetch_add(&next_chunk, 1, __ATOMIC_RELAXED);
        if (t >= nchunks) break;
        chunk_rec* r = &R[t];
        const int64_t p0 = (est * t) / nchunks, p1 = (est * (t + 1)) / nchunks;
        /* the words this chunk may read: its range of the stream and a sample's worth beyond */
        int64_t want = (*pos + 2 * (p1 + 4096)) / MT_N + 1;
        if (want > nblocks) want = nblocks;
        int64_t got;
        while ((got = __atomic_load_n(&ext_blocks, __ATOMIC_ACQUIRE)) < want)
            CPU_RELAX();
        beta_ctx L = C;                                /* never reads past what is published */
        const int64_t avail = ((got + 1) * MT_N - *pos) / 2;
        if (avail < L.cap) L.cap = avail;
        if (!r->m0) {
            r->cap = (int64_t)((double)(p1 - p0) / per * 1.25) + 1024;
            r->m0 = (int64_t*)malloc((size_t)r->cap * sizeof(int64_t));
            r->src0 = (int64_t*)malloc((size_t)r->cap * sizeof(int64_t));
            r->val = (double*)malloc((size_t)r->cap * sizeof(double));
        }
        r->n = 0;
        r->rc = 0;
        if (!r->m0 || !r->src0 || !r->val) { r->rc = -1; continue; }
        walk_state W = entry;
        if (t > 0) {                                   /* the guess: a sample starts here */
            W.m = p0 + (((p0 & 1) != cls) ? 1 : 0);
            W.src = SRC_NONE;
            W.cache = 0.0;
        }
        while (W.m < p1 && r->n < r->cap && r->n < n) {
            r->m0[r->n] = W.m;
            r->src0[r->n] = W.src;
            if (walk_beta(&L, &W, &r->val[r->n])) { r->rc = -2; break; }
            r->n++;
        }
        r->end = W;
    }
    }
    if (first > 0) goto stitch_resume;

    t_walk = now_s();
    /* ---- stitch the chunks in order */
stitch_resume:
    for (; t < nchunks && count < n && rc == 0; t++) {
        chunk_rec* r = &R[t];
        if (r->rc == -1) { rc = -1; break; }
        const int64_t p1 = (est * (t + 1)) / nchunks;
        int64_t k = 0;
        int merged = 0;
        while (count < n) {
            while (k < r->n && r->m0[k] < S.m) k++;
            if (k < r->n && r->m0[k] == S.m && r->src0[k] == S.src) { merged = 1; break; }
            if (S.m >= p1) break;                    /* no coincidence inside this chunk */
            if (walk_beta(&C, &S, &out[count])) { rc = -2; break; }
            count++;
        }
        if (rc) continue;
        if (!merged) {
            const int now = (int)((S.m + (S.src != SRC_NONE)) & 1);
            if (now != cls && t + 1 < nchunks && count < n) {   /* the class flipped (see above) */
                cls = now;
                first = ++t;
                goto walk_again;
            }
            continue;
        }
        int64_t take = r->n - k;
        /* a chunk whose walk stopped early (record full / stream end) is used up to there */
        if (take > n - count) take = n - count;
        memcpy(out + count, r->val + k, (size_t)take * sizeof(double));
        count += take;
        if (k + take < r->n) {                       /* stopped inside the chunk: n reached */
            S.m = r->m0[k + take];
            S.src = r->src0[k + take];
            S.cache = 0.0;
            if (S.src >= 0) {
                const double x1 = 2.0 * dbl_at(C.w, S.src) - 1.0;
                const double x2 = 2.0 * dbl_at(C.w, S.src + 1) - 1.0;
                const double r2 = x1 * x1 + x2 * x2;
                S.cache = sqrt(-2.0 * log(r2) / r2) * x1;
            } else if (S.src == SRC_ENTRY) {
                S.cache = entry.cache;
            }
        } else {
            S = r->end;
        }
    }
    while (rc == 0 && count < n) {                   /* the estimate fell short: finish in line */
        if (walk_beta(&C, &S, &out[count])) { rc = -2; break; }
        count++;
    }
    if (rc == 0) {
        *has_gauss = S.src != SRC_NONE;
        *gauss = S.src != SRC_NONE ? S.cache : 0.0;
        int64_t cursor = *pos + 2 * S.m;
        int64_t blk = cursor / MT_N, off = cursor % MT_N;
        if (off == 0 && blk > 0) { blk -= 1; off = MT_N; }
        memcpy(key, raw + blk * MT_N, MT_N * sizeof(uint32_t));
        *pos = (int32_t)off;
    }
    if (trace)
        fprintf(stderr, "beta n=%lld threads=%d chunks=%d: pilot %.2f extend %.2f walk %.2f "
                "stitch %.2f ms (%.2f doubles/sample)\n", (long long)n, nthreads, nchunks,
                (t_pilot - t_start) * 1e3, (t_extend - t_pilot) * 1e3, (t_walk - t_extend) * 1e3,
                (now_s() - t_walk) * 1e3, per);
    for (int t = 0; t < nchunks; t++) { free(R[t].m0); free(R[t].src0); free(R[t].val); }
    free(R);
    return rc;
}
