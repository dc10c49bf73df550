// Per-sample transit / eclipse model evaluated inside the sm_100a kernels.
//
// Follows the semantics fixed by oracle/quadmodel.py (the restatement of
// pytransit==2.2 QuadraticModel(interpolate=False), which the reference calls at
// triceratops/likelihoods.py:348-349, :414-415, :421-422):
//   orbit        mean anomaly -> true anomaly through the bilinear (e, M) table, then the
//                projected separation z
//   occultation  Mandel & Agol (2002) quadratic limb darkening with the Hastings K/E
//                polynomials and Bulirsch's iteration for the third-kind integral
//   supersample  mean over nsamples sub-exposures
//
// All functions are __host__ __device__ so the identical source can be exercised on the CPU by
// tests/hostcheck (logic check of the window / branch structure without a GPU); the product
// path only ever runs them on the device.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define TRI_HD __host__ __device__ __forceinline__
#else
#define TRI_HD inline
#endif

namespace tri {

constexpr double kPi = 3.14159265358979323846;
constexpr double kHalfPi = 0.5 * kPi;
constexpr double kTwoPi = 2.0 * kPi;
constexpr double kInvPi = 1.0 / kPi;
constexpr double kInvTwoPi = 1.0 / kTwoPi;

// astropy >= 4.0 constants, cgs (CODATA 2018 / IAU 2015 nominal); reference reads them at
// likelihoods.py:17-21 and marginal_likelihoods.py:13-17.
constexpr double kG = 6.6743e-08;
constexpr double kMsun = 1.988409870698051e+33;
constexpr double kRsun = 69570000000.0;
constexpr double kRearth = 637810000.0;

constexpr int kTableNe = 256;
constexpr int kTableNm = 512;
constexpr double kTableMaxE = 0.95;

// (e, M) -> f - M table, resident in HBM (1 MiB, served from L1/L2).
struct OrbitTable {
    const double* tae;  // [kTableNe][kTableNm]
    double de, dm;
    double inv_dm;
};

// Folded light curve, time-sorted, resident in HBM (optionally staged in shared memory).
struct LightCurve {
    const double* time;    // [npts] ascending
    const double* flux;    // [npts]
    const double* prefix;  // [npts+1] running sum of (flux-1)^2, prefix[0]=0
    int npts;
    int nsamples;
    double sigma;
    double exptime;
    double tmin, tmax;
};

// Everything z(t) needs for one sample.
struct Orbit {
    double k;          // radius ratio
    double e;          // eccentricity
    double n_rate;     // 2 pi / P
    double c0;         // time at which the mean anomaly is zero: -(M_tr * P / 2 pi)
    double ma_tr;      // mean anomaly at mid-transit
    double sinw, cosw; // of the model's periastron angle w = (90 - argp) pi/180 (+ pi for secondary)
    double sini2;      // sin^2 i
    double a1me2;      // a/R* (1 - e^2)
    double ae;         // table blend weight in e
    const double* row0;
    const double* row1;
    bool table_clamped;  // e beyond the table: last cell extrapolated (monotonicity not guaranteed)
};

TRI_HD double mean_anomaly_offset(double e, double w) {
    double s, c;
    sincos(kHalfPi - w, &s, &c);
    double off = atan2(sqrt(1.0 - e * e) * s, e + c);
    off -= e * sin(off);
    return off;
}

TRI_HD void orbit_setup(Orbit& o, const OrbitTable& T, double k, double p, double a_rs,
                        double inc_rad, double e, double w) {
    o.k = k;
    o.e = e;
    o.n_rate = kTwoPi / p;
    double off = mean_anomaly_offset(e, w);
    o.ma_tr = off;
    o.c0 = 0.0 - off * p / kTwoPi;
    sincos(w, &o.sinw, &o.cosw);
    double si = sin(inc_rad);
    o.sini2 = si * si;
    o.a1me2 = a_rs * (1.0 - e * e);
    int ie = (int)floor(e / T.de);
    o.table_clamped = false;
    if (ie > kTableNe - 2) {
        ie = kTableNe - 2;
        o.table_clamped = true;
    }
    o.ae = (e - T.de * ie) / T.de;
    o.row0 = T.tae + (size_t)ie * kTableNm;
    o.row1 = o.row0 + kTableNm;
}

// table part of the true anomaly: f(M) for M already reduced to [0, 2 pi)
TRI_HD double ta_from_ma(const Orbit& o, const OrbitTable& T, double ma) {
    double x, s;
    if (ma < kPi) { x = ma; s = 1.0; } else { x = kTwoPi - ma; s = -1.0; }
    // the blend is continuous across cells, so x*inv_dm landing one cell off the oracle's
    // floor(x/dm) at a cell edge only changes the result by rounding
    int im = (int)floor(x * T.inv_dm);
    if (im > kTableNm - 2) im = kTableNm - 2;
    double am = (x - im * T.dm) * T.inv_dm;
#if defined(__CUDA_ARCH__)
    double t00 = __ldg(o.row0 + im), t01 = __ldg(o.row0 + im + 1);
    double t10 = __ldg(o.row1 + im), t11 = __ldg(o.row1 + im + 1);
#else
    double t00 = o.row0[im], t01 = o.row0[im + 1], t10 = o.row1[im], t11 = o.row1[im + 1];
#endif
    double d = t00 * (1.0 - o.ae) * (1.0 - am) + t10 * o.ae * (1.0 - am)
             + t01 * (1.0 - o.ae) * am + t11 * o.ae * am;
    return ma + s * d;
}

TRI_HD double mean_anomaly(const Orbit& o, double t) {
    double x = (t - o.c0) * o.n_rate;
    double ma = fma(-kTwoPi, floor(x * kInvTwoPi), x);
    if (ma < 0.0) ma += kTwoPi;
    if (ma >= kTwoPi) ma -= kTwoPi;
    return ma;
}

TRI_HD double z_from_ta(const Orbit& o, double ta) {
    double st, ct;
    sincos(ta, &st, &ct);
    double swt = o.sinw * ct + o.cosw * st;  // sin(w + f)
    double z = o.a1me2 / (1.0 + o.e * ct) * sqrt(1.0 - swt * swt * o.sini2);
    return swt < 0.0 ? -z : z;
}

TRI_HD double z_at(const Orbit& o, const OrbitTable& T, double t) {
    return z_from_ta(o, ta_from_ma(o, T, mean_anomaly(o, t)));
}

// ---------------------------------------------------------------- elliptic integrals
// K and E share m1 = 1 - q^2 and its logarithm (Hastings, A&S 17.3.34 / 17.3.36).
TRI_HD void ellke(double q, double& Kk, double& Ek) {
    double m1 = 1.0 - q * q;
    double lg = log(m1);
    double ek1 = 1.38629436112 + m1 * (0.09666344259 + m1 * (0.03590092383
               + m1 * (0.03742563713 + m1 * 0.01451196212)));
    double ek2 = 0.5 + m1 * (0.12498593597 + m1 * (0.06880248576
               + m1 * (0.03328355346 + m1 * 0.00441787012)));
    double ee1 = 1.0 + m1 * (0.44325141463 + m1 * (0.0626060122
               + m1 * (0.04757383546 + m1 * 0.01736506451)));
    double ee2 = m1 * (0.2499836831 + m1 * (0.09200180037 + m1 * (0.04069697526
               + m1 * 0.00526449639)));
    Kk = ek1 - ek2 * lg;
    Ek = ee1 - ee2 * lg;
}

// Bulirsch (1965) third-kind integral.  One reciprocal per sweep instead of two divisions, and
// the convergence test |1 - kc/g| > 1e-8 written without its division; the sweep count can only
// differ from the oracle's when the test is within rounding of its threshold, where one more
// (quadratically convergent) sweep changes the value below 1e-16.
TRI_HD double ellpicb(double n, double q) {
    double kc = sqrt(1.0 - q * q);
    double e = kc;
    double p = sqrt(n + 1.0);
    double m0 = 1.0, c = 1.0;
    double d = 1.0 / p;
    for (int it = 0; it < 64; ++it) {
        double ip = 1.0 / p;
        double f = c;
        c = fma(d, ip, c);
        double g = e * ip;
        d = 2.0 * fma(f, g, d);
        p = g + p;
        g = m0;
        m0 = kc + m0;
        if (fabs(g - kc) > 1e-8 * g) {
            kc = 2.0 * sqrt(e);
            e = kc * m0;
        } else {
            return kHalfPi * fma(c, m0, d) / (m0 * (m0 + p));
        }
    }
    return 0.0;
}

// ---------------------------------------------------------------------- occultation
// Limb-darkening mix for one sample: flux = 1 - (c_le*le + c_ld*ld + u2*ed) * inv_omega
struct Limb {
    double c_le, c_ld, u2, inv_omega;
};

TRI_HD void limb_setup(Limb& L, double u1, double u2) {
    L.c_le = 1.0 - u1 - 2.0 * u2;
    L.c_ld = u1 + 2.0 * u2;
    L.u2 = u2;
    L.inv_omega = 1.0 / (1.0 - u1 / 3.0 - u2 / 6.0);
}

// Relative flux at separation z (any sign), radius ratio k.  Mirrors oracle eval_quad():
// same case tests, same expressions; K, E and Pi are evaluated once for whichever of the two
// general cases (limb-crossing III / interior IV) applies so that divergent lanes share them.
TRI_HD double occult_quad(double z, double k, const Limb& L) {
    if (fabs(z - k) < 1e-6) z += 1e-6;
    if (z > 1.0 + k || z < 0.0) return 1.0;
    if (k >= 1.0 && z <= k - 1.0) return 0.0;

    const double k2 = k * k, z2 = z * z;
    const double x1 = (k - z) * (k - z), x2 = (k + z) * (k + z), x3 = k * k - z * z;
    double le = 0.0, ld = 0.0, ed = 0.0, kap0 = 0.0, kap1 = 0.0;

    const bool partial = (z >= fabs(1.0 - k) && z <= 1.0 + k);
    if (partial) {
        kap1 = acos(fmin((1.0 - k2 + z2) * 0.5 / z, 1.0));
        kap0 = acos(fmin((k2 + z2 - 1.0) * 0.5 / k / z, 1.0));
        double t = 1.0 + z2 - k2;
        le = (k2 * kap0 + kap1 - 0.5 * sqrt(fmax(4.0 * z2 - t * t, 0.0))) * kInvPi;
    }
    if (z <= 1.0 - k) le = k2;

    const bool edge = fabs(z - k) < 1e-4 * (z + k);
    const bool case3 = !edge && ((z > 0.5 + fabs(k - 0.5) && z < 1.0 + k)
                                 || (k > 0.5 && z > fabs(1.0 - k) * 1.0001 && z < k));
    const bool case4 = !edge && !case3 && (k <= 1.0 && z < (1.0 - k) * 1.0001);

    if (case3 || case4) {
        double q, n;
        if (case3) {
            q = sqrt((1.0 - x1) * 0.25 / z / k);
            n = 1.0 / x1 - 1.0;
        } else {
            q = sqrt((x2 - x1) / (1.0 - x1));
            n = x2 / x1 - 1.0;
        }
        double Kk, Ek;
        ellke(q, Kk, Ek);
        double Pk = ellpicb(n, q);
        double pterm = 3.0 * x3 / x1 * Pk;
        if (case3) {
            ld = 1.0 / 9.0 * kInvPi / sqrt(k * z)
               * (((1.0 - x2) * (2.0 * x2 + x1 - 3.0) - 3.0 * x3 * (x2 - 2.0)) * Kk
                  + 4.0 * k * z * (z2 + 7.0 * k2 - 4.0) * Ek - pterm);
            if (z < k) ld += 2.0 / 3.0;
            ed = 0.5 * kInvPi * (kap1 + k2 * (k2 + 2.0 * z2) * kap0
               - (1.0 + 5.0 * k2 + z2) * 0.25 * sqrt((1.0 - x1) * (x2 - 1.0)));
        } else {
            ld = 2.0 / 9.0 * kInvPi / sqrt(1.0 - x1)
               * ((1.0 - 5.0 * z2 + k2 + x3 * x3) * Kk
                  + (1.0 - x1) * (z2 + 7.0 * k2 - 4.0) * Ek - pterm);
            if (z < k) ld += 2.0 / 3.0;
            if (fabs(k + z - 1.0) < 1e-4)
                ld = 2.0 / 3.0 * kInvPi * acos(1.0 - 2.0 * k)
                   - 4.0 / 9.0 * kInvPi * sqrt(k * (1.0 - k)) * (3.0 + 2.0 * k - 8.0 * k2);
            ed = k2 * 0.5 * (k2 + 2.0 * z2);
        }
    } else if (edge) {
        if (k == 0.5) {
            ld = 1.0 / 3.0 - 4.0 * kInvPi / 9.0;
            ed = 3.0 / 32.0;
        } else if (z > 0.5) {
            double Kk, Ek;
            ellke(0.5 / k, Kk, Ek);
            ld = 1.0 / 3.0 + 16.0 * k / 9.0 * kInvPi * (2.0 * k2 - 1.0) * Ek
               - (32.0 * (k2 * k2) - 20.0 * k2 + 3.0) / 9.0 * kInvPi / k * Kk;
            ed = 0.5 * kInvPi * (kap1 + k2 * (k2 + 2.0 * z2) * kap0
               - (1.0 + 5.0 * k2 + z2) * 0.25 * sqrt((1.0 - x1) * (x2 - 1.0)));
        } else {
            double Kk, Ek;
            ellke(2.0 * k, Kk, Ek);
            ld = 1.0 / 3.0 + 2.0 / 9.0 * kInvPi * (4.0 * (2.0 * k2 - 1.0) * Ek
               + (1.0 - 4.0 * k2) * Kk);
            ed = k2 * 0.5 * (k2 + 2.0 * z2);
        }
    }
    return 1.0 - (L.c_le * le + L.c_ld * ld + L.u2 * ed) * L.inv_omega;
}

// Supersampled flux at one time stamp (mean over sub-exposures).
TRI_HD double model_point(const Orbit& o, const OrbitTable& T, const Limb& L, double t,
                          double exptime, int ns) {
    double acc = 0.0;
    const double inv_ns = 1.0 / ns;
    for (int is = 1; is <= ns; ++is) {
        double toff = exptime * ((is - 0.5) * inv_ns - 0.5);
        double z = z_at(o, T, t + toff);
        acc += (z > 1.0 + o.k) ? 1.0 : occult_quad(z, o.k, L);
    }
    return acc / ns;
}

// -------------------------------------------------------------------- transit window
// Conservative time interval outside of which every sub-exposure has z > 1 + k or z < 0, so the
// model is exactly 1 there (and stays exactly 1 after dilution).  In transit requires
// |cos(w + f)| <= (1 + k) / r with r >= a(1 - e), i.e. f within delta of the transit anomaly;
// the edges in mean anomaly are found by bisection on the SAME interpolated f(M) the model
// uses, so table error cannot move a point across the edge.  Returns false when no window can
// be guaranteed (caller then evaluates every point).
struct Window {
    double t_lo, t_hi;
};

TRI_HD bool in_arc(const Orbit& o, double ta, double cmax) {
    double st, ct;
    sincos(ta, &st, &ct);
    double swt = o.sinw * ct + o.cosw * st;
    double cwt = o.cosw * ct - o.sinw * st;
    return swt > 0.0 && fabs(cwt) <= cmax;
}

TRI_HD bool transit_window(const Orbit& o, const OrbitTable& T, double a_rs, double p,
                           const LightCurve& lc, Window& win) {
    if (o.table_clamped) return false;
    double rmin = a_rs * (1.0 - o.e);
    double cmax = (1.0 + o.k) / rmin * (1.0 + 1e-9) + 1e-12;
    if (!(cmax < 0.95)) return false;  // wide arcs: not worth it / not safe
    // f(M) is monotone (bilinear blend of monotone rows); M is taken relative to mid-transit.
    double ma0 = o.ma_tr - kTwoPi * floor(o.ma_tr * kInvTwoPi);
    if (!in_arc(o, ta_from_ma(o, T, ma0 >= kTwoPi ? ma0 - kTwoPi : ma0), cmax)) return false;
    double edge[2];
    for (int side = 0; side < 2; ++side) {
        double sgn = side ? 1.0 : -1.0;
        double lo = 0.0, hi = kPi;  // offset from mid-transit; lo inside the arc, hi outside
        {
            double m = ma0 + sgn * hi;
            m -= kTwoPi * floor(m * kInvTwoPi);
            if (m >= kTwoPi) m -= kTwoPi;
            if (in_arc(o, ta_from_ma(o, T, m), cmax)) return false;
        }
        for (int it = 0; it < 26; ++it) {
            double mid = 0.5 * (lo + hi);
            double m = ma0 + sgn * mid;
            m -= kTwoPi * floor(m * kInvTwoPi);
            if (m >= kTwoPi) m -= kTwoPi;
            if (m < 0.0) m += kTwoPi;
            if (in_arc(o, ta_from_ma(o, T, m), cmax)) lo = mid; else hi = mid;
        }
        edge[side] = hi;  // first offset known to be outside
    }
    // offsets in mean anomaly -> time (mid-transit is t = 0); pad for rounding
    double pad = 1e-9 * p + 1e-12;
    win.t_lo = -edge[0] / o.n_rate - pad;
    win.t_hi = edge[1] / o.n_rate + pad;
    // other images of the window (t +- P) must not reach the light curve
    double half = 0.5 * lc.exptime;
    if (lc.tmax + half >= win.t_lo + p) return false;
    if (lc.tmin - half <= win.t_hi - p) return false;
    return true;
}

// first index with time[j] >= x
TRI_HD int lower_bound(const double* t, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (t[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// -------------------------------------------------------------------------- samples
// Dilution of the host-only model (likelihoods.py:352-357 and :427-438):
//   m -> (m + d1)/(1 + d1)   [EB only]   then   m -> (m + d2)/(1 + d2)
struct Dilution {
    double d1, d2;
    bool two_stage;
};

TRI_HD double dilute(const Dilution& D, double m) {
    if (D.two_stage) m = (m + D.d1) / (1.0 + D.d1);
    return (m + D.d2) / (1.0 + D.d2);
}

}  // namespace tri
