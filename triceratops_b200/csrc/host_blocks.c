/*
 * The element-wise preparation of a scenario's prior draws (everything an lnZ_* function does
 * between its last draw from numpy's generator and the submission of its columns to the GPU:
 * inverse-CDF transforms, stellar relations, flux ratios, companion / background priors, limb
 * darkening look-ups), for one scenario in one call, without the GIL, chunk by chunk over the
 * host threads.  Host-side preparation only: nothing of the light-curve path lives here.
 *
 * It mirrors the numpy statements of triceratops_b200/marginal_likelihoods.py, priors.py,
 * funcs.py and _ldc.py (which restate reference marginal_likelihoods.py:39-2362, priors.py and
 * funcs.py) operation by operation, in the same order and association, compiled without
 * contraction, so every arithmetic result has the same bits.  The transcendental functions are
 * not re-implemented: the caller passes numpy's own compiled inner loops (taken from the ufunc
 * objects numpy.power / log10 / log / exp / arccos), so those values are computed by the very
 * code numpy would run.  tests/test_host_blocks.py holds this file to bit-equality with the
 * numpy path for every scenario and option; the Python side also checks it once per process
 * and keeps using numpy if the check fails.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef TB_PROFILE        /* section timers (single-threaded runs only): gcc -DTB_PROFILE */
#include <stdio.h>
#include <x86intrin.h>
enum { S_PIECE, S_INC, S_STELLAR, S_FLUX, S_INTERP, S_BOUND, S_BG, S_LDC, S_TAKE, S_ECC, S_N };
static const char* s_names[S_N] = {"piecewise", "inc", "stellar", "fluxratio", "interp",
                                   "bound_rest", "background_rest", "ldc", "take", "ecc"};
static unsigned long long s_acc[S_N];
#define TIC unsigned long long _t0 = __rdtsc()
#define TOC(k) s_acc[k] += __rdtsc() - _t0
void trih_block_profile_dump(double per) {
    for (int i = 0; i < S_N; i++) {
        fprintf(stderr, "%-16s %8.1f\n", s_names[i], (double)s_acc[i] / per);
        s_acc[i] = 0;
    }
}
#else
#define TIC ((void)0)
#define TOC(k) ((void)0)
#endif

#define CH 8192                 /* elements per chunk: ~25 live arrays of 64 KB stay in L2 */
#define NBUF 28

typedef void (*np_loop)(char**, const intptr_t*, const intptr_t*, void*);
typedef struct { np_loop fn; void* data; } np_fn;

typedef struct { const double* t; const double* c; int32_t n; int32_t k; } tb_spline;

typedef struct {
    int32_t nseg;               /* 0: every element becomes `fill` */
    int32_t pad;
    double fill, norm;
    double knot[3], integ[3], p1[3], amp[3], e0[3], inv[3];
} tb_powerlaw;

typedef struct {
    int32_t mode;               /* 0: zeros (MOLUSC table); 1: planet scenarios; 2: EB scenarios */
    int32_t m_ge_1;             /* M_act >= 1.0 */
    double d;                   /* 1000 / plx */
    const double* xp;           /* contrasts */
    const double* fp;           /* separations */
    int64_t nxp;
    double K1, au, f1, f2, f3, slope, slope2, two_f1, half_alpha, alpha_dlogP;
    double t2, t23, t234, t2345, t4, t45, M_act;
} tb_bound;

typedef struct {
    int32_t mode;               /* 0: no contrast curve (constant); 1: contrast curve */
    int32_t pad;
    double constant, K;
    const double* xp;
    const double* fp;
    int64_t nxp;
} tb_background;

typedef struct {
    const int64_t* code;        /* sorted node codes Teff * 100 + round(logg * 10) */
    const double* u1;
    const double* u2;
    int64_t n;
    double cap;
} tb_ldc;

typedef struct {
    int32_t kind, flatpriors, has_cc, molusc;
    int64_t N;
    int32_t nthreads, pad;
    np_fn f_pow, f_log10, f_log, f_exp, f_arccos;
    const double *x_rp, *x_inc, *x_q, *x_w, *c_comp;
    double* x_e;
    const int64_t* idxs;
    double M_s, R_s, Teff, c_lo, inc_norm, ecc_exp, G, Msun, Rsun, rp_flat_A;
    tb_powerlaw rp_hi, rp_lo, q, q_comp;
    tb_spline hot_R, hot_T, cool_R, cool_T, flux_tess, flux_cc;
    double f0_tess, f0_cc;
    tb_bound bound;
    tb_background bgp;
    tb_ldc ldc;
    int64_t ntab;
    const double *bg_mass, *bg_radius, *bg_teff, *bg_logg, *bg_fr, *bg_band, *bg_fr_tess,
        *bg_fr_cc, *bg_u1, *bg_u2;
    double* out[16];
    uint8_t* extra;
    /* contrast tables that are not non-decreasing (numpy.interp then depends on the order of
     * the queries, see stitch_interp): the signed magnitude differences and the interval found
     * for every draw, N each; NULL otherwise */
    double* interp_delta;
    int32_t* interp_j;
} tb_args;

int64_t trih_block_args_size(void) { return (int64_t)sizeof(tb_args); }

enum { K_TTP, K_TEB, K_PTP, K_PEB, K_STP, K_SEB, K_DTP, K_DEB, K_BTP, K_BEB };

/* ---- numpy's loops ----------------------------------------------------------------------- */
static void v_unary(const np_fn* f, const double* in, double* out, int64_t m) {
    char* args[2] = {(char*)in, (char*)out};
    intptr_t dims[1] = {(intptr_t)m}, steps[2] = {8, 8};
    f->fn(args, dims, steps, f->data);
}

/* numpy.power(x, e): the ufunc itself */
static void v_power_vs(const np_fn* f, const double* x, double e, double* out, int64_t m) {
    char* args[3] = {(char*)x, (char*)&e, (char*)out};
    intptr_t dims[1] = {(intptr_t)m}, steps[3] = {8, 0, 8};
    f->fn(args, dims, steps, f->data);
}

/* b ** y with a scalar base: numpy.power(b, y) */
static void v_power_sv(const np_fn* f, double b, const double* y, double* out, int64_t m) {
    char* args[3] = {(char*)&b, (char*)y, (char*)out};
    intptr_t dims[1] = {(intptr_t)m}, steps[3] = {0, 8, 8};
    f->fn(args, dims, steps, f->data);
}

/* x ** e as the ndarray operator evaluates it for a Python scalar exponent: numpy's
 * fast_scalar_power turns 1, 2, -1, 0, 0.5 into positive, square, reciprocal, ones, sqrt */
static void v_pow_op(const np_fn* f, const double* x, double e, double* out, int64_t m) {
    if (e == 1.0) {
        if (out != x) memcpy(out, x, (size_t)m * sizeof(double));
    } else if (e == 2.0) {
        for (int64_t i = 0; i < m; i++) out[i] = x[i] * x[i];
    } else if (e == -1.0) {
        for (int64_t i = 0; i < m; i++) out[i] = 1.0 / x[i];
    } else if (e == 0.0) {
        for (int64_t i = 0; i < m; i++) out[i] = 1.0;
    } else if (e == 0.5) {
        for (int64_t i = 0; i < m; i++) out[i] = sqrt(x[i]);
    } else {
        v_power_vs(f, x, e, out, m);
    }
}

/* ---- FITPACK splev (same recurrence as csrc/host_prep.c) ------------------------------------ */
/* One de Boor step: f = hh / (T[li] - T[lj]) (or the zero-width rule), the two updates. */
#define DEBOOR(hq, hq1, hh, li, lj)                          \
    do {                                                     \
        const double tli = T[li], tlj = T[lj];               \
        if (tli == tlj) {                                    \
            hq1 = 0.0;                                       \
        } else {                                             \
            const double f = (hh) / (tli - tlj);             \
            hq = hq + f * (tli - arg);                       \
            hq1 = f * (arg - tlj);                           \
        }                                                    \
    } while (0)

/* cubic splines (every relation of funcs.py is one): the generic recurrence below written out
 * for k = 3, operation for operation, and the knot interval found by counting instead of by
 * bisection (the knots are few and sorted: the same interval, without unpredictable branches) */
static inline double splev_cubic(const tb_spline* s, double arg) {
    const int nk1 = s->n - 4;
    const double* T = s->t - 1;
    const double* C = s->c - 1;
    int l = 4;
    for (int i = 5; i <= nk1; i++) l += (arg >= T[i]);
    double h1 = 1.0, h2 = 0.0, h3 = 0.0, h4 = 0.0, a, b, c;
    a = h1; h1 = 0.0;
    DEBOOR(h1, h2, a, l + 1, l);
    a = h1; b = h2; h1 = 0.0;
    DEBOOR(h1, h2, a, l + 1, l - 1);
    DEBOOR(h2, h3, b, l + 2, l);
    a = h1; b = h2; c = h3; h1 = 0.0;
    DEBOOR(h1, h2, a, l + 1, l - 2);
    DEBOOR(h2, h3, b, l + 2, l - 1);
    DEBOOR(h3, h4, c, l + 3, l);
    double sp = 0.0;
    sp = sp + C[l - 3] * h1;
    sp = sp + C[l - 2] * h2;
    sp = sp + C[l - 1] * h3;
    sp = sp + C[l] * h4;
    return sp;
}

static inline double splev1(const tb_spline* s, double arg) {
    if (s->k == 3) return splev_cubic(s, arg);
    const int k = s->k, k1 = k + 1, nk1 = s->n - k1;
    const double* T = s->t - 1;
    const double* C = s->c - 1;
    int lo = k1, hi = nk1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (arg >= T[mid]) lo = mid; else hi = mid - 1;
    }
    const int l = lo;
    double h[7], hh[6];
    h[1] = 1.0;
    for (int j = 1; j <= k; j++) {
        for (int q = 1; q <= j; q++) hh[q] = h[q];
        h[1] = 0.0;
        for (int q = 1; q <= j; q++) {
            const int li = l + q, lj = li - j;
            if (T[li] == T[lj]) {
                h[q + 1] = 0.0;
            } else {
                const double f = hh[q] / (T[li] - T[lj]);
                h[q] = h[q] + f * (T[li] - arg);
                h[q + 1] = f * (arg - T[lj]);
            }
        }
    }
    double sp = 0.0;
    int ll = l - k1;
    for (int j = 1; j <= k1; j++) {
        ll = ll + 1;
        sp = sp + C[ll] * h[j];
    }
    return sp;
}

/* ---- numpy.interp (numpy/_core/src/multiarray/compiled_base.c: arr_interp) ------------------ */
#define LIKELY_IN_CACHE_SIZE 8
static int64_t search_with_guess(double key, const double* arr, int64_t len, int64_t guess) {
    int64_t imin = 0, imax = len;
    if (key > arr[len - 1]) return len;
    if (key < arr[0]) return -1;
    if (len <= 4) {
        int64_t i;
        for (i = 1; i < len && key >= arr[i]; ++i) {}
        return i - 1;
    }
    if (guess > len - 3) guess = len - 3;
    if (guess < 1) guess = 1;
    if (key < arr[guess]) {
        if (key < arr[guess - 1]) {
            imax = guess - 1;
            if (guess > LIKELY_IN_CACHE_SIZE && key >= arr[guess - LIKELY_IN_CACHE_SIZE])
                imin = guess - LIKELY_IN_CACHE_SIZE;
        } else {
            return guess - 1;
        }
    } else {
        if (key < arr[guess + 1]) return guess;
        if (key < arr[guess + 2]) return guess + 1;
        imin = guess + 2;
        if (guess < len - LIKELY_IN_CACHE_SIZE - 1 && key < arr[guess + LIKELY_IN_CACHE_SIZE])
            imax = guess + LIKELY_IN_CACHE_SIZE;
    }
    while (imin < imax) {
        const int64_t imid = imin + ((imax - imin) >> 1);
        if (key >= arr[imid]) imin = imid + 1; else imax = imid;
    }
    return imin - 1;
}

/* Keys for which the search cannot depend on where it starts: every comparison `key >= arr[i]`
 * is then true up to some index and false after it, as on a sorted table, and any search finds
 * the same interval.  A key k is unsafe only if an out-of-order pair i < i', arr[i] > arr[i']
 * has arr[i'] <= k < arr[i]; [*lo, *hi) covers all of those (empty when the table is sorted). */
static void unsafe_keys(const double* arr, int64_t len, double* lo, double* hi) {
    *lo = INFINITY;
    *hi = -INFINITY;
    for (int64_t i = 0; i < len; i++) {
        if (isnan(arr[i])) { *lo = -INFINITY; *hi = INFINITY; return; }
        for (int64_t k = i + 1; k < len; k++) {
            if (arr[i] > arr[k]) {
                if (arr[k] < *lo) *lo = arr[k];
                if (arr[i] > *hi) *hi = arr[i];
            }
        }
    }
}

/* the interval numpy's search returns for a safe key: the last index with arr[idx] <= key,
 * -1 below the table, len above it */
static inline int64_t search_sorted(double key, const double* arr, int64_t len) {
    if (key > arr[len - 1]) return len;
    if (key < arr[0]) return -1;
    int64_t lo = 0, n = len;
    while (n > 1) {
        const int64_t half = n >> 1;
        lo = (key >= arr[lo + half]) ? lo + half : lo;
        n -= half;
    }
    return lo;
}

static inline double interp_at(double xv, int64_t j, const double* xp, const double* fp,
                               int64_t nxp) {
    if (j == -1) return fp[0];
    if (j == nxp) return fp[nxp - 1];
    if (j == nxp - 1) return fp[j];
    if (xp[j] == xv) return fp[j];
    const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
    double r = slope * (xv - xp[j]) + fp[j];
    if (isnan(r)) {
        r = slope * (xv - xp[j + 1]) + fp[j + 1];
        if (isnan(r) && fp[j] == fp[j + 1]) r = fp[j];
    }
    return r;
}

/* jrec (optional): the search state after every element, numpy's `j` */
static void v_interp(const double* x, const double* xp, const double* fp, int64_t nxp,
                     double* out, int64_t m, int32_t* jrec) {
    const double lval = fp[0], rval = fp[nxp - 1];
    if (nxp == 1) {
        const double xv = xp[0], fv = fp[0];
        for (int64_t i = 0; i < m; i++)
            out[i] = (x[i] < xv) ? lval : ((x[i] > xv) ? rval : fv);
        return;
    }
    /* (numpy tabulates the slopes when nxp <= m: the same quotient either way) */
    double ulo, uhi;
    unsafe_keys(xp, nxp, &ulo, &uhi);
    int64_t j = 0;
    for (int64_t i = 0; i < m; i++) {
        const double xv = x[i];
        if (isnan(xv)) {
            out[i] = xv;
        } else {
            /* numpy's guess-based search only where its starting point could matter */
            j = (xv >= ulo && xv < uhi) ? search_with_guess(xv, xp, nxp, j)
                                        : search_sorted(xv, xp, nxp);
            out[i] = interp_at(xv, j, xp, fp, nxp);
        }
        if (jrec) jrec[i] = (int32_t)j;
    }
}

/* ---- per-thread scratch --------------------------------------------------------------------- */
typedef struct {
    double* b[NBUF];
    int32_t* ix;
    uint8_t* sel;
    int32_t* perm;        /* splev_many: draws grouped by knot interval */
    int32_t* subset;      /* v_stellar: the draws of one mass branch */
    uint8_t* l8;          /* splev_many: knot interval of every draw */
} scratch;

#if defined(__GNUC__) && defined(__x86_64__)
#define CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define CLONES
#endif

/* The cubic de Boor recurrence of splev_cubic for draws that share the knot interval l: the
 * knots and coefficients are loop constants, the loop body is straight-line IEEE arithmetic on
 * the argument (the same operations in the same order, so the same bits) and vectorises,
 * divisions included.  Requires the six knot differences to be non-zero (the caller checks). */
CLONES static void deboor3_run(const double* T, const double* C, int l, const double* xin,
                               double* xout, int64_t n) {
    const double tm2 = T[l - 2], tm1 = T[l - 1], t0 = T[l], t1 = T[l + 1], t2 = T[l + 2],
                 t3 = T[l + 3];
    const double c0 = C[l - 3], c1 = C[l - 2], c2 = C[l - 1], c3 = C[l];
    const double f1 = 1.0 / (t1 - t0);
    const double d21 = t1 - tm1, d22 = t2 - t0, d31 = t1 - tm2, d32 = t2 - tm1, d33 = t3 - t0;
    for (int64_t q = 0; q < n; q++) {
        const double arg = xin[q];
        double h1, h2, h3, h4, a, b, c, f;
        h1 = 0.0 + f1 * (t1 - arg);
        h2 = f1 * (arg - t0);
        a = h1; b = h2;
        f = a / d21;
        h1 = 0.0 + f * (t1 - arg);
        h2 = f * (arg - tm1);
        f = b / d22;
        h2 = h2 + f * (t2 - arg);
        h3 = f * (arg - t0);
        a = h1; b = h2; c = h3;
        f = a / d31;
        h1 = 0.0 + f * (t1 - arg);
        h2 = f * (arg - tm2);
        f = b / d32;
        h2 = h2 + f * (t2 - arg);
        h3 = f * (arg - tm1);
        f = c / d33;
        h3 = h3 + f * (t3 - arg);
        h4 = f * (arg - t0);
        double sp = 0.0;
        sp = sp + c0 * h1;
        sp = sp + c1 * h2;
        sp = sp + c2 * h3;
        sp = sp + c3 * h4;
        xout[q] = sp;
    }
}

/* out[i] = spline(x[i]) for the draws i = sel[0..n) (sel == NULL: i = 0..n).  Cubic splines:
 * the draws are grouped by knot interval (counting sort) and every group runs through
 * deboor3_run; anything else goes draw by draw through splev1. */
static void splev_many(const tb_spline* s, const double* x, const int32_t* sel, int64_t n,
                       double* out, scratch* W) {
    if (n <= 0) return;
    if (s->k != 3 || s->n - 4 > 250) {
        for (int64_t q = 0; q < n; q++) {
            const int64_t i = sel ? sel[q] : q;
            out[i] = splev1(s, x[i]);
        }
        return;
    }
    const int nk1 = s->n - 4;
    const double* T = s->t - 1;
    const double* C = s->c - 1;
    double* tin = W->b[NBUF - 2];
    double* tout = W->b[NBUF - 3];
    double* arg = W->b[NBUF - 4];
    uint8_t* l8 = W->l8;
    if (sel) {
        for (int64_t q = 0; q < n; q++) arg[q] = x[sel[q]];
    } else {
        memcpy(arg, x, (size_t)n * sizeof(double));
    }
    /* knot interval of every draw: 4 + the number of interior knots at or below it */
    for (int64_t q = 0; q < n; q++) l8[q] = 4;
    for (int i = 5; i <= nk1; i++) {
        const double knot = T[i];
        for (int64_t q = 0; q < n; q++) l8[q] += (uint8_t)(arg[q] >= knot);
    }
    int32_t count[256] = {0}, start[256] = {0}, fill[256];
    for (int64_t q = 0; q < n; q++) count[l8[q]]++;
    int32_t acc = 0;
    for (int l = 4; l <= nk1; l++) { start[l] = acc; acc += count[l]; }
    memcpy(fill, start, sizeof(fill));
    for (int64_t q = 0; q < n; q++) {
        const int32_t pos = fill[l8[q]]++;
        W->perm[pos] = (int32_t)q;
        tin[pos] = arg[q];
    }
    for (int l = 4; l <= nk1; l++) {
        const int32_t c = count[l];
        if (!c) continue;
        const int64_t o = start[l];
        const int flat = T[l + 1] == T[l] || T[l + 1] == T[l - 1] || T[l + 2] == T[l]
                         || T[l + 1] == T[l - 2] || T[l + 2] == T[l - 1] || T[l + 3] == T[l];
        if (flat) {
            for (int64_t q = o; q < o + c; q++) tout[q] = splev_cubic(s, tin[q]);
        } else {
            deboor3_run(T, C, l, tin + o, tout + o, c);
        }
    }
    if (sel) {
        for (int64_t q = 0; q < n; q++) out[sel[W->perm[q]]] = tout[q];
    } else {
        for (int64_t q = 0; q < n; q++) out[W->perm[q]] = tout[q];
    }
}

/* ---- priors.py: _piecewise_powerlaw ---------------------------------------------------------- */
static void v_piecewise(const tb_args* A, const tb_powerlaw* S, const double* x, const uint8_t* sel,
                        double* out, int64_t m, scratch* W) {
    TIC;
    double* tmp = W->b[NBUF - 1];
    int32_t* ix = W->ix;
    for (int k = 0; k < S->nseg; k++) {
        int64_t c = 0;
        for (int64_t i = 0; i < m; i++) {
            int in = x[i] <= S->knot[k];
            if (k > 0) in = (x[i] > S->knot[k - 1]) & in;
            if (sel) in = in & sel[i];
            ix[c] = (int32_t)i;                 /* (branch-free compaction) */
            c += in;
        }
        if (!c) continue;
        for (int64_t q = 0; q < c; q++) {
            double u = x[ix[q]] / S->norm;
            for (int j = 0; j < k; j++) u = u - S->integ[j];
            u = u * S->p1[k];
            u = u / S->amp[k];
            tmp[q] = u + S->e0[k];
        }
        v_pow_op(&A->f_pow, tmp, S->inv[k], tmp, c);
        for (int64_t q = 0; q < c; q++) out[ix[q]] = tmp[q];
    }
    TOC(S_PIECE);
}

/* sample_q / sample_q_companion: the deviates become mass ratios (or np.full(n, 1.0)) */
static void v_mass_ratio(const tb_args* A, const tb_powerlaw* S, const double* x, double* out,
                         int64_t m, scratch* W) {
    if (S->nseg == 0) {
        for (int64_t i = 0; i < m; i++) out[i] = S->fill;
        return;
    }
    memcpy(out, x, (size_t)m * sizeof(double));           /* (out never aliases x here) */
    v_piecewise(A, S, x, NULL, out, m, W);
}

/* sample_rp: host masses per draw (hm) or one value */
static void v_planet_radius(const tb_args* A, const double* x, const double* hm, double hm_s,
                            double* out, int64_t m, scratch* W) {
    if (A->flatpriors) {
        for (int64_t i = 0; i < m; i++) out[i] = x[i] / A->rp_flat_A + 0.5;
        return;
    }
    memcpy(out, x, (size_t)m * sizeof(double));
    uint8_t* sel = W->sel;
    for (int64_t i = 0; i < m; i++) sel[i] = (hm ? hm[i] : hm_s) > 0.45;
    v_piecewise(A, &A->rp_hi, x, sel, out, m, W);
    for (int64_t i = 0; i < m; i++) sel[i] = (hm ? hm[i] : hm_s) <= 0.45;
    v_piecewise(A, &A->rp_lo, x, sel, out, m, W);
}

/* sample_inc: arccos(c_lo - x / norm) * 180 / pi */
static void v_inc(const tb_args* A, const double* x, double* out, int64_t m) {
    TIC;
    for (int64_t i = 0; i < m; i++) out[i] = A->c_lo - x[i] / A->inc_norm;
    v_unary(&A->f_arccos, out, out, m);
    for (int64_t i = 0; i < m; i++) out[i] = out[i] * 180.0 / 3.141592653589793;
    TOC(S_INC);
}

static void v_argp(const double* x, double* out, int64_t m) {
    for (int64_t i = 0; i < m; i++) out[i] = x[i] * 360.0;
}

/* funcs.stellar_relations: caps per draw (maxR / maxT) or one value each */
static void v_stellar(const tb_args* A, const double* mass, const double* maxR, double maxR_s,
                      const double* maxT, double maxT_s, double* R, double* T, int64_t m,
                      scratch* W) {
    TIC;
    /* Radii[hot] = spline(m[hot]) ...: the two mass branches, each as one batch */
    for (int64_t i = 0; i < m; i++) { R[i] = 0.0; if (T) T[i] = 0.0; }
    for (int branch = 0; branch < 2; branch++) {
        int64_t c = 0;
        for (int64_t i = 0; i < m; i++) {
            const int in = branch == 0 ? mass[i] > 0.63 : mass[i] <= 0.63;
            W->subset[c] = (int32_t)i;
            c += in;
        }
        if (!c) continue;
        splev_many(branch == 0 ? &A->hot_R : &A->cool_R, mass, W->subset, c, R, W);
        if (T) splev_many(branch == 0 ? &A->hot_T : &A->cool_T, mass, W->subset, c, T, W);
    }
    for (int64_t i = 0; i < m; i++) {
        double r = R[i];
        const double cr = maxR ? maxR[i] : maxR_s;
        if (r > cr) r = cr;
        if (r < 0.1) r = 0.1;
        R[i] = r;
        if (T) {
            double t = T[i];
            const double ct = maxT ? maxT[i] : maxT_s;
            if (t > ct) t = ct;
            if (t < 2800.0) t = 2800.0;
            T[i] = t;
        }
    }
    TOC(S_STELLAR);
}

/* marginal_likelihoods._fluxratio: f / (f + f0) with f = 10 ** spline(mass) */
static void v_fluxratio(const tb_args* A, const tb_spline* s, double f0, const double* mass,
                        double* out, int64_t m, scratch* W) {
    TIC;
    splev_many(s, mass, NULL, m, out, W);
    v_power_sv(&A->f_pow, 10.0, out, out, m);
    for (int64_t i = 0; i < m; i++) out[i] = out[i] / (out[i] + f0);
    TOC(S_FLUX);
}

static inline void v_odds(const double* fr, double* out, int64_t m) {      /* fr / (1 - fr) */
    for (int64_t i = 0; i < m; i++) out[i] = fr[i] / (1.0 - fr[i]);
}

static void v_clip(double* lnprior, const double* delta, int64_t m) {
    for (int64_t i = 0; i < m; i++) {
        if (lnprior[i] > 0.0) lnprior[i] = 0.0;
        if (delta[i] > 0.0) lnprior[i] = -INFINITY;
    }
}

/* marginal_likelihoods._bound_prior over priors._bound_companion_lnprior; `term` is the flux-ratio
 * term (fr_tess or cc_term()).  b0..b3: scratch. */
static void bound_from_separation(const tb_args* A, double* lp, const double* delta,
                                  double* lnprior, int64_t m, double* ex);

static void v_bound_prior(const tb_args* A, int64_t lo, const double* term, double* lnprior,
                          int64_t m, double* delta, double* lp, double* ex, double* b3) {
    const tb_bound* B = &A->bound;
    if (B->mode == 0) {
        for (int64_t i = 0; i < m; i++) lnprior[i] = 0.0;
        return;
    }
    v_unary(&A->f_log10, term, delta, m);
    for (int64_t i = 0; i < m; i++) { delta[i] = 2.5 * delta[i]; b3[i] = fabs(delta[i]); }
    if (A->interp_j) memcpy(A->interp_delta + lo, delta, (size_t)m * sizeof(double));
    {
        TIC;
        v_interp(b3, B->xp, B->fp, B->nxp, lp, m,          /* separation_at_contrast */
                 A->interp_j ? A->interp_j + lo : NULL);
        TOC(S_INTERP);
    }
    bound_from_separation(A, lp, delta, lnprior, m, ex);
}

/* the rest of the bound-companion prior, from the separations [arcsec] in lp (overwritten) */
static void bound_from_separation(const tb_args* A, double* lp, const double* delta,
                                  double* lnprior, int64_t m, double* ex) {
    TIC;
    const tb_bound* B = &A->bound;
    for (int64_t i = 0; i < m; i++) lp[i] = (B->d * lp[i]) * B->au;       /* seps * au */
    v_power_vs(&A->f_pow, lp, 3.0, lp, m);                                 /* ** 3 */
    for (int64_t i = 0; i < m; i++) lp[i] = sqrt(B->K1 * lp[i]) / 86400.0;
    v_unary(&A->f_log10, lp, lp, m);
    for (int64_t i = 0; i < m; i++) ex[i] = -0.3 * lp[i];
    v_unary(&A->f_exp, ex, ex, m);
    const int first = B->mode == 2;
    for (int64_t i = 0; i < m; i++) {
        const double x = lp[i];
        double f = 0.0;
        if (x >= 1.0 && x < 2.0) {
            if (first) f = (0.5 * (x - 1.0)) * (B->two_f1 + B->slope * (x - 1.0));
        } else if (x >= 2.0 && x < 3.4) {
            if (first)
                f = B->t2 + (B->half_alpha * ((x * x - 5.4 * x) + 6.8) + B->f2 * (x - 2.0));
        } else if (x >= 3.4 && x < 5.5) {
            const double t4p = (B->alpha_dlogP * (x - 3.4) + B->f2 * (x - 3.4))
                               + B->slope2 * ((0.238095 * (x * x) - 0.952381 * x) + 0.485714);
            f = first ? B->t23 + t4p : t4p;
        } else if (x >= 5.5 && x < 8.0) {
            const double t5p = B->f3 * (3.33333 - 17.3566 * ex[i]);
            f = first ? B->t234 + t5p : B->t4 + t5p;
        } else if (x >= 8.0) {
            f = first ? B->t2345 : B->t45;
        }
        if (!B->m_ge_1) {
            f = 0.65 * f + (0.35 * f) * B->M_act;
            if (f < 0.0) f = 0.0;
        }
        lnprior[i] = f;
    }
    v_unary(&A->f_log, lnprior, lnprior, m);
    v_clip(lnprior, delta, m);
    TOC(S_BOUND);
}

/* marginal_likelihoods._background_prior: dmag is dmag_tess (no contrast curve) or dmag_cc */
static void background_from_separation(const tb_args* A, double* lnprior, const double* dmag,
                                       int64_t m) {
    TIC;
    for (int64_t i = 0; i < m; i++) lnprior[i] = A->bgp.K * (lnprior[i] * lnprior[i]);
    v_unary(&A->f_log, lnprior, lnprior, m);
    v_clip(lnprior, dmag, m);
    TOC(S_BG);
}

static void v_background_prior(const tb_args* A, int64_t lo, const double* dmag, double* lnprior,
                               int64_t m, double* b0) {
    const tb_background* B = &A->bgp;
    if (B->mode == 0) {
        for (int64_t i = 0; i < m; i++) lnprior[i] = B->constant;
        v_clip(lnprior, dmag, m);
        return;
    }
    for (int64_t i = 0; i < m; i++) b0[i] = fabs(dmag[i]);
    if (A->interp_j) memcpy(A->interp_delta + lo, dmag, (size_t)m * sizeof(double));
    {
        TIC;
        v_interp(b0, B->xp, B->fp, B->nxp, lnprior, m, A->interp_j ? A->interp_j + lo : NULL);
        TOC(S_INTERP);
    }
    background_from_separation(A, lnprior, dmag, m);
}

/* _ldc.LdcGrid.at_Z_rounded */
static int v_ldc(const tb_args* A, const double* teff, const double* logg, double* u1, double* u2,
                 int64_t m) {
    const tb_ldc* L = &A->ldc;
    for (int64_t i = 0; i < m; i++) {
        double rg = rint(logg[i] / 0.5) * 0.5;
        if (rg < 3.5) rg = 3.5;
        if (rg > 5.0) rg = 5.0;
        double rT = rint(teff[i] / 250.0) * 250.0;
        if (rT < 3500.0) rT = 3500.0;
        if (rT > L->cap) rT = L->cap;
        if (!(rT == rint(rT)) || !(fabs(rT) < 1e15)) return -4;
        const int64_t code = (int64_t)rT * 100 + (int64_t)rint(rg * 10.0);
        int64_t lo = 0, hi = L->n;                      /* searchsorted, side = left */
        while (lo < hi) {
            const int64_t mid = lo + ((hi - lo) >> 1);
            if (L->code[mid] < code) lo = mid + 1; else hi = mid;
        }
        if (lo >= L->n || L->code[lo] != code) return -4;
        u1[i] = L->u1[lo];
        u2[i] = L->u2[lo];
    }
    return 0;
}

static int v_take(const double* tab, int64_t ntab, const int64_t* idx, double* out, int64_t m) {
    for (int64_t i = 0; i < m; i++) {
        const int64_t j = idx[i];
        if (j < 0 || j >= ntab) return -5;
        out[i] = tab[j];
    }
    return 0;
}

/* _companion_stars: properties of the drawn bound companions when they host the event */
static int v_companion_stars(const tb_args* A, const double* qs_comp, double* masses_comp,
                             double* radii_comp, double* teffs_comp, double* frc, double* u1,
                             double* u2, int64_t m, double* b0, scratch* W) {
    for (int64_t i = 0; i < m; i++) masses_comp[i] = qs_comp[i] * A->M_s;
    v_stellar(A, masses_comp, NULL, A->R_s, NULL, A->Teff, radii_comp, teffs_comp, m, W);
    for (int64_t i = 0; i < m; i++) {
        const double c = radii_comp[i] * A->Rsun;
        b0[i] = (A->G * (masses_comp[i] * A->Msun)) / (c * c);
    }
    v_unary(&A->f_log10, b0, b0, m);
    v_fluxratio(A, &A->flux_tess, A->f0_tess, masses_comp, frc, m, W);
    return v_ldc(A, teffs_comp, b0, u1, u2, m);
}

#define OUT(k) (A->out[k] + lo)

static int run_chunk(const tb_args* A, int64_t lo, int64_t m, scratch* W) {
    double** b = W->b;
    const double* x_rp = A->x_rp ? A->x_rp + lo : NULL;
    const double* x_inc = A->x_inc + lo;
    const double* x_q = A->x_q ? A->x_q + lo : NULL;
    const double* x_w = A->x_w + lo;
    const double* c_comp = A->c_comp ? A->c_comp + lo : NULL;
    const int64_t* idxs = A->idxs ? A->idxs + lo : NULL;
    uint8_t* extra = A->extra ? A->extra + lo : NULL;
    int rc = 0;
    const int binary = A->kind == K_TEB || A->kind == K_PEB || A->kind == K_SEB
                       || A->kind == K_DEB || A->kind == K_BEB;
    if (binary) {
        /* np.power(x_e, 1 / a, out = x_e): scipy.stats.powerlaw.rvs from its deviates */
        TIC;
        v_power_vs(&A->f_pow, A->x_e + lo, A->ecc_exp, A->x_e + lo, m);
        TOC(S_ECC);
    }
    double* qs_comp = b[0];
    if (c_comp) {
        if (A->molusc) memcpy(qs_comp, c_comp, (size_t)m * sizeof(double));
        else v_mass_ratio(A, &A->q_comp, c_comp, qs_comp, m, W);
        if (extra) for (int64_t i = 0; i < m; i++) extra[i] = qs_comp[i] != 0.0;
    }
    switch (A->kind) {
    case K_TTP:
        v_planet_radius(A, x_rp, NULL, A->M_s, OUT(0), m, W);
        v_inc(A, x_inc, OUT(1), m);
        v_argp(x_w, OUT(2), m);
        break;
    case K_TEB:
    case K_PEB:
    case K_DEB: {
        double *incs = OUT(0), *qs = OUT(1), *argps = OUT(2), *masses = OUT(3), *radii = OUT(4),
               *fr = OUT(5), *mtot = OUT(6);
        v_inc(A, x_inc, incs, m);
        v_mass_ratio(A, &A->q, x_q, qs, m, W);
        v_argp(x_w, argps, m);
        for (int64_t i = 0; i < m; i++) masses[i] = qs[i] * A->M_s;
        v_stellar(A, masses, NULL, A->R_s, NULL, A->Teff, radii, NULL, m, W);
        v_fluxratio(A, &A->flux_tess, A->f0_tess, masses, fr, m, W);
        for (int64_t i = 0; i < m; i++) mtot[i] = A->M_s + masses[i];
        if (A->kind == K_PEB) {
            double *frc = OUT(7), *lnprior = OUT(8);
            for (int64_t i = 0; i < m; i++) b[1][i] = qs_comp[i] * A->M_s;        /* masses_comp */
            v_fluxratio(A, &A->flux_tess, A->f0_tess, b[1], frc, m, W);
            if (A->has_cc) {
                v_fluxratio(A, &A->flux_cc, A->f0_cc, b[1], b[2], m, W);
                v_odds(b[2], b[2], m);
            } else {
                v_odds(frc, b[2], m);
            }
            v_bound_prior(A, lo, b[2], lnprior, m, b[3], b[4], b[5], b[6]);
        } else if (A->kind == K_DEB) {
            double *cfr = OUT(7), *lnprior = OUT(8);
            if ((rc = v_take(A->bg_fr, A->ntab, idxs, cfr, m))) return rc;
            if (A->has_cc) {
                if ((rc = v_take(A->bg_band, A->ntab, idxs, b[2], m))) return rc;
            } else {
                v_odds(cfr, b[2], m);
                v_unary(&A->f_log10, b[2], b[2], m);
                for (int64_t i = 0; i < m; i++) b[2][i] = 2.5 * b[2][i];
            }
            v_background_prior(A, lo, b[2], lnprior, m, b[3]);
        }
        break;
    }
    case K_PTP:
    case K_DTP: {
        double *rps = OUT(0), *incs = OUT(1), *argps = OUT(2);
        v_planet_radius(A, x_rp, NULL, A->M_s, rps, m, W);
        v_inc(A, x_inc, incs, m);
        v_argp(x_w, argps, m);
        if (A->kind == K_PTP) {
            double *frc = OUT(3), *lnprior = OUT(4);
            for (int64_t i = 0; i < m; i++) b[1][i] = qs_comp[i] * A->M_s;
            v_fluxratio(A, &A->flux_tess, A->f0_tess, b[1], frc, m, W);
            if (A->has_cc) {
                v_fluxratio(A, &A->flux_cc, A->f0_cc, b[1], b[2], m, W);
                v_odds(b[2], b[2], m);
            } else {
                v_odds(frc, b[2], m);
            }
            v_bound_prior(A, lo, b[2], lnprior, m, b[3], b[4], b[5], b[6]);
        } else {
            double *cfr = OUT(3), *lnprior = OUT(4);
            if ((rc = v_take(A->bg_fr, A->ntab, idxs, cfr, m))) return rc;
            if (A->has_cc) {
                if ((rc = v_take(A->bg_band, A->ntab, idxs, b[2], m))) return rc;
            } else {
                v_odds(cfr, b[2], m);
                v_unary(&A->f_log10, b[2], b[2], m);
                for (int64_t i = 0; i < m; i++) b[2][i] = 2.5 * b[2][i];
            }
            v_background_prior(A, lo, b[2], lnprior, m, b[3]);
        }
        break;
    }
    case K_STP: {
        double *rps = OUT(0), *incs = OUT(1), *argps = OUT(2), *masses_comp = OUT(3),
               *radii_comp = OUT(4), *frc = OUT(5), *u1 = OUT(6), *u2 = OUT(7), *lnprior = OUT(8);
        for (int64_t i = 0; i < m; i++) b[1][i] = qs_comp[i] * A->M_s;            /* host masses */
        v_planet_radius(A, x_rp, b[1], 0.0, rps, m, W);
        v_inc(A, x_inc, incs, m);
        v_argp(x_w, argps, m);
        if ((rc = v_companion_stars(A, qs_comp, masses_comp, radii_comp, b[2], frc, u1, u2, m,
                                    b[3], W)))
            return rc;
        if (A->has_cc) {
            v_fluxratio(A, &A->flux_cc, A->f0_cc, masses_comp, b[4], m, W);
            v_odds(b[4], b[4], m);
        } else {
            v_odds(frc, b[4], m);
        }
        v_bound_prior(A, lo, b[4], lnprior, m, b[5], b[6], b[7], b[8]);
        break;
    }
    case K_SEB: {
        double *incs = OUT(0), *qs = OUT(1), *argps = OUT(2), *masses_comp = OUT(3),
               *radii_comp = OUT(4), *u1 = OUT(5), *u2 = OUT(6), *mtot = OUT(7), *masses = OUT(8),
               *radii = OUT(9), *fr = OUT(10), *frc = OUT(11), *lnprior = OUT(12);
        v_inc(A, x_inc, incs, m);
        v_mass_ratio(A, &A->q, x_q, qs, m, W);
        v_argp(x_w, argps, m);
        double* teffs_comp = b[2];
        if ((rc = v_companion_stars(A, qs_comp, masses_comp, radii_comp, teffs_comp, frc, u1, u2,
                                    m, b[3], W)))
            return rc;
        for (int64_t i = 0; i < m; i++) masses[i] = qs[i] * masses_comp[i];
        v_stellar(A, masses, radii_comp, 0.0, teffs_comp, 0.0, radii, NULL, m, W);
        v_fluxratio(A, &A->flux_tess, A->f0_tess, masses, fr, m, W);
        if (A->has_cc) {
            v_fluxratio(A, &A->flux_cc, A->f0_cc, masses, b[4], m, W);
            v_fluxratio(A, &A->flux_cc, A->f0_cc, masses_comp, b[5], m, W);
            v_odds(b[4], b[4], m);
            v_odds(b[5], b[5], m);
        } else {
            v_odds(fr, b[4], m);
            v_odds(frc, b[5], m);
        }
        for (int64_t i = 0; i < m; i++) b[4][i] = b[5][i] + b[4][i];
        v_bound_prior(A, lo, b[4], lnprior, m, b[5], b[6], b[7], b[8]);
        for (int64_t i = 0; i < m; i++) mtot[i] = masses_comp[i] + masses[i];
        break;
    }
    case K_BTP: {
        double *hm = OUT(0), *rps = OUT(1), *incs = OUT(2), *argps = OUT(3), *cfr = OUT(4),
               *lnprior = OUT(5), *hr = OUT(6), *u1 = OUT(7), *u2 = OUT(8);
        if ((rc = v_take(A->bg_mass, A->ntab, idxs, hm, m))) return rc;
        v_planet_radius(A, x_rp, hm, 0.0, rps, m, W);
        v_inc(A, x_inc, incs, m);
        v_argp(x_w, argps, m);
        v_take(A->bg_fr, A->ntab, idxs, cfr, m);
        if (A->has_cc) {
            v_take(A->bg_band, A->ntab, idxs, b[2], m);
        } else {
            v_odds(cfr, b[2], m);
            v_unary(&A->f_log10, b[2], b[2], m);
            for (int64_t i = 0; i < m; i++) b[2][i] = 2.5 * b[2][i];
        }
        v_background_prior(A, lo, b[2], lnprior, m, b[3]);
        for (int64_t i = 0; i < m; i++)
            extra[i] = (A->bg_logg[idxs[i]] >= 3.5) & (A->bg_teff[idxs[i]] <= 10000.0);
        v_take(A->bg_radius, A->ntab, idxs, hr, m);
        v_take(A->bg_u1, A->ntab, idxs, u1, m);
        v_take(A->bg_u2, A->ntab, idxs, u2, m);
        break;
    }
    case K_BEB: {
        double *incs = OUT(0), *qs = OUT(1), *argps = OUT(2), *hm = OUT(3), *hr = OUT(4),
               *u1 = OUT(5), *u2 = OUT(6), *mtot = OUT(7), *masses = OUT(8), *radii = OUT(9),
               *fr = OUT(10), *cfr = OUT(11), *lnprior = OUT(12);
        v_inc(A, x_inc, incs, m);
        v_mass_ratio(A, &A->q, x_q, qs, m, W);
        v_argp(x_w, argps, m);
        if ((rc = v_take(A->bg_mass, A->ntab, idxs, hm, m))) return rc;
        v_take(A->bg_radius, A->ntab, idxs, hr, m);
        v_take(A->bg_fr, A->ntab, idxs, cfr, m);
        for (int64_t i = 0; i < m; i++) masses[i] = qs[i] * hm[i];
        v_take(A->bg_teff, A->ntab, idxs, b[1], m);
        v_stellar(A, masses, hr, 0.0, b[1], 0.0, radii, NULL, m, W);
        /* distance_corrected("TESS"): _fluxratio(masses) * (cfr_band / _fluxratio(host_masses)) */
        v_take(A->bg_fr_tess, A->ntab, idxs, b[2], m);
        v_fluxratio(A, &A->flux_tess, A->f0_tess, hm, b[3], m, W);
        v_fluxratio(A, &A->flux_tess, A->f0_tess, masses, fr, m, W);
        for (int64_t i = 0; i < m; i++) fr[i] = fr[i] * (b[2][i] / b[3][i]);
        if (!A->has_cc) {
            v_odds(cfr, b[4], m);
            v_odds(fr, b[5], m);
        } else {
            v_take(A->bg_fr_cc, A->ntab, idxs, b[2], m);                      /* cfr_cc */
            v_fluxratio(A, &A->flux_cc, A->f0_cc, hm, b[3], m, W);
            v_fluxratio(A, &A->flux_cc, A->f0_cc, masses, b[6], m, W);
            for (int64_t i = 0; i < m; i++) b[6][i] = b[6][i] * (b[2][i] / b[3][i]);   /* fr_cc */
            v_odds(b[2], b[4], m);
            v_odds(b[6], b[5], m);
        }
        for (int64_t i = 0; i < m; i++) b[4][i] = b[4][i] + b[5][i];
        v_unary(&A->f_log10, b[4], b[4], m);
        for (int64_t i = 0; i < m; i++) b[4][i] = 2.5 * b[4][i];                 /* dmag */
        v_background_prior(A, lo, b[4], lnprior, m, b[5]);
        for (int64_t i = 0; i < m; i++)
            extra[i] = (A->bg_logg[idxs[i]] >= 3.5) & (A->bg_teff[idxs[i]] <= 10000.0);
        v_take(A->bg_u1, A->ntab, idxs, u1, m);
        v_take(A->bg_u2, A->ntab, idxs, u2, m);
        for (int64_t i = 0; i < m; i++) mtot[i] = hm[i] + masses[i];
        break;
    }
    default:
        return -9;
    }
    return 0;
}

/* numpy.interp searches the table from where the previous query ended.  On a table that is not
 * non-decreasing the interval it finds can depend on that starting point, i.e. on the order of
 * the queries -- the reference evaluates all N in one call.  The chunks above each started
 * their search afresh; this pass walks the chunk boundaries in order with the true search
 * state and re-evaluates the draws whose interval comes out different (as soon as a draw's
 * state coincides with the recorded one, the rest of its chunk stands). */
static void stitch_interp(const tb_args* A) {
    static const int lnprior_col[] = {-1, -1, 4, 8, 8, 12, 4, 8, 5, 12};
    const int bound = A->kind == K_PTP || A->kind == K_PEB || A->kind == K_STP
                      || A->kind == K_SEB;
    const double* xp = bound ? A->bound.xp : A->bgp.xp;
    const double* fp = bound ? A->bound.fp : A->bgp.fp;
    const int64_t nxp = bound ? A->bound.nxp : A->bgp.nxp;
    double* lnprior = A->out[lnprior_col[A->kind]];
    int32_t* jrec = A->interp_j;
    int64_t g = 0;
    double ulo, uhi;
    unsafe_keys(xp, nxp, &ulo, &uhi);
    for (int64_t lo = 0; lo < A->N; lo += CH) {
        const int64_t hi = lo + CH < A->N ? lo + CH : A->N;
        for (int64_t i = lo; lo > 0 && i < hi; i++) {
            const double delta = A->interp_delta[i], xv = fabs(delta);
            const int64_t jt = isnan(xv) ? g
                               : (xv >= ulo && xv < uhi) ? search_with_guess(xv, xp, nxp, g)
                                                         : search_sorted(xv, xp, nxp);
            if (jt == jrec[i]) break;
            if (!isnan(xv)) {
                double sep = interp_at(xv, jt, xp, fp, nxp), ex;
                if (bound) bound_from_separation(A, &sep, &delta, &lnprior[i], 1, &ex);
                else { lnprior[i] = sep; background_from_separation(A, &lnprior[i], &delta, 1); }
            }
            jrec[i] = (int32_t)jt;
            g = jt;
        }
        g = jrec[hi - 1];
    }
}

int trih_scenario_block(const tb_args* A) {
    const int64_t N = A->N;
    if (N <= 0) return 0;
    const int64_t nchunks = (N + CH - 1) / CH;
    int nt = A->nthreads < 1 ? 1 : A->nthreads;
    if ((int64_t)nt > nchunks) nt = (int)nchunks;
    int err = 0;
#pragma omp parallel num_threads(nt)
    {
        scratch W;
        double* arena = (double*)malloc((size_t)NBUF * CH * sizeof(double));
        W.ix = (int32_t*)malloc((size_t)CH * sizeof(int32_t));
        W.sel = (uint8_t*)malloc((size_t)CH);
        W.perm = (int32_t*)malloc((size_t)CH * sizeof(int32_t));
        W.subset = (int32_t*)malloc((size_t)(CH + 1) * sizeof(int32_t));
        W.l8 = (uint8_t*)malloc((size_t)CH);
        int ok = arena && W.ix && W.sel && W.perm && W.subset && W.l8;
        for (int i = 0; i < NBUF; i++) W.b[i] = ok ? arena + (size_t)i * CH : NULL;
#pragma omp for schedule(dynamic, 1)
        for (int64_t c = 0; c < nchunks; c++) {
            int rc;
            if (!ok) {
                rc = -1;
            } else {
                const int64_t lo = c * CH, m = lo + CH <= N ? CH : N - lo;
                rc = run_chunk(A, lo, m, &W);
            }
            if (rc) {
#pragma omp atomic write
                err = rc;
            }
        }
        free(arena);
        free(W.ix);
        free(W.sel);
        free(W.perm);
        free(W.subset);
        free(W.l8);
    }
    if (!err && A->interp_j) stitch_interp(A);
    return err;
}
