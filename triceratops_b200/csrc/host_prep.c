/*
 * Host-side helpers for the PRIOR-DRAW preparation of the lnZ_* functions (not a compute
 * fallback: nothing of the light-curve path lives here).
 *
 * trih_splev: evaluation of a FITPACK B-spline (t, c, k) at m points with extrapolation, the
 * operation behind scipy's InterpolatedUnivariateSpline.__call__ that the reference uses for its
 * stellar relations (funcs.py:31-119).  scipy evaluates it single-threaded under the GIL and it
 * is ~45 % of the host time of a calc_probs at N = 1e6.  This is the same de Boor recurrence in
 * the same operation order (Dierckx, fpbspl/splev), so the results are bit-identical, spread
 * over the host cores with OpenMP.  Build: gcc -O2 -ffp-contract=off -fopenmp.
 */
#include <stdint.h>

#define KMAX 5

int trih_splev(const double* t, int n, const double* c, int k, const double* x, double* y,
               int64_t m, int nthreads) {
    if (k < 1 || k > KMAX || n < 2 * (k + 1)) return -1;
    const int k1 = k + 1;
    const int nk1 = n - k1;
    /* 1-based knot/coefficient access as in the Fortran original */
    const double* T = t - 1;
    const double* C = c - 1;
    /* the thread count is the caller's, per call: the process-wide OpenMP setting (torch, BLAS,
     * the OMP_NUM_THREADS=1 that torchrun exports) is left alone */
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t i = 0; i < m; i++) {
        const double arg = x[i];
        /* knot interval T(l) <= arg < T(l+1), clamped to [k1, nk1] (extrapolation uses the end
         * polynomial pieces) */
        int lo = k1, hi = nk1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (arg >= T[mid]) lo = mid; else hi = mid - 1;
        }
        const int l = lo;
        double h[KMAX + 2], hh[KMAX + 1];
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
        y[i] = sp;
    }
    return 0;
}

/* out[i] = tab[idx[i]]: the gather behind `property_of_background_star[idxs]` (reference
 * marginal_likelihoods.py:1463-1480 and alike, ~30 of them per calc_probs).  numpy's fancy
 * indexing does this single-threaded while holding the GIL, which stalls the thread that is
 * drawing the next scenario's priors; here it runs without the GIL over the caller's threads.
 * Returns -1 (nothing written past the first bad element's chunk) if an index is out of range. */
int trih_take_f64(const double* tab, int64_t ntab, const int64_t* idx, double* out, int64_t n,
                  int nthreads) {
    if (nthreads < 1) nthreads = 1;
    int bad = 0;
#pragma omp parallel for schedule(static) num_threads(nthreads) reduction(| : bad)
    for (int64_t i = 0; i < n; i++) {
        int64_t j = idx[i];
        if (j < 0) j += ntab;                  /* numpy's negative indices */
        if (j < 0 || j >= ntab) { bad |= 1; continue; }
        out[i] = tab[j];
    }
    return bad ? -1 : 0;
}
