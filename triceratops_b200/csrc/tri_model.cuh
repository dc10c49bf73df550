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

// Reciprocal, reciprocal square root and square root for operands that are known to be normal
// and well scaled (radius ratios, separations, AGM iterates): hardware seed (MUFU.RCP64H /
// RSQ64H, relative error e0 <= 2^-20) + ONE third-order correction step, without the IEEE slow
// paths (subnormals, correct rounding):
//   1/x       y (1 + e + e^2),             e = 1 - x y       residual e^3   (3 FP64 ops)
//   1/sqrt x  y (1 + e/2 + 3 e^2/8),       e = 1 - x y^2     residual 5e^3/16 (5 FP64 ops)
// i.e. ~2^-60 before rounding, ~1 ulp after it: far inside the 1e-9 tolerance on lnL (the
// second-order Newton pairs they replace cost 4 and 8 ops); on the host (tests/hostcheck) they
// are the exact operations.
TRI_HD double fast_rcp(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    e = fma(e, e, e);
    return fma(y, e, y);
#else
    return 1.0 / x;
#endif
}

TRI_HD double fast_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-(x * y), y, 1.0);
    double p = fma(0.375, e, 0.5) * e;
    return fma(y, p, y);
#else
    return 1.0 / sqrt(x);
#endif
}

// sqrt(x) for x > 0 (no zero / negative handling: AGM iterates)
TRI_HD double fast_sqrt_pos(double x) {
#if defined(__CUDA_ARCH__)
    return x * fast_rsqrt(x);
#else
    return sqrt(x);
#endif
}

// sqrt(x) for x >= 0 (0 -> 0, negative -> NaN as with sqrt)
TRI_HD double fast_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    double s = x * fast_rsqrt(x);
    return x == 0.0 ? 0.0 : s;
#else
    return sqrt(x);
#endif
}

// sqrt(max(x, 0)): 0 and the slightly negative arguments that rounding can leave in
// 1 - sin^2(..) sin^2 i give 0 (x rsqrt(x) is NaN there and fmax returns its other operand)
TRI_HD double fast_sqrt_clamped(double x) {
#if defined(__CUDA_ARCH__)
    return fmax(x * fast_rsqrt(x), 0.0);
#else
    return x > 0.0 ? sqrt(x) : 0.0;
#endif
}

// high 32 bits of a double (sign, exponent, 20 mantissa bits): ordered like the value for
// non-negative doubles
TRI_HD unsigned hi_word(double x) {
#if defined(__CUDA_ARCH__)
    return (unsigned)__double2hiint(x);
#else
    union { double d; unsigned long long u; } cv;
    cv.d = x;
    return (unsigned)(cv.u >> 32);
#endif
}

// ---- polynomial tables in constant memory ------------------------------------------------------
// Literal double constants cost two uniform-register moves each on sm_100a; adjacent entries of
// a __constant__ table are fetched two at a time (LDCU.128).  The hot loop evaluates ~50
// coefficients per model point, so the tables below (and hand-written sincos / log kernels
// that use them instead of the CUDA library routines' embedded literals) remove ~10 % of the
// issued instructions.  On the host (tests/hostcheck) the same tables are plain arrays.
#define TRI_TABLE(name, n, ...)                                  \
    static const double name##_host[n] = {__VA_ARGS__};           \
    TRI_DEVICE_TABLE(name, n, __VA_ARGS__)
#if defined(__CUDACC__)
#define TRI_DEVICE_TABLE(name, n, ...) __constant__ double name##_dev[n] = {__VA_ARGS__};
#else
#define TRI_DEVICE_TABLE(name, n, ...)
#endif
#if defined(__CUDA_ARCH__)
#define TRI_T(name) name##_dev
#else
#define TRI_T(name) name##_host
#endif

// sin/cos on [-pi/4, pi/4] (fdlibm __kernel_sin / __kernel_cos minimax coefficients) and the
// two-term Cody-Waite split of pi/2
TRI_TABLE(kTrig, 16,
          6.36619772367581382433e-01,   /* 0  2/pi      */
          1.57079632679489655800e+00,   /* 1  pi/2 high */
          6.12323399573676603587e-17,   /* 2  pi/2 low  */
          -1.66666666666666324348e-01,  /* 3  S1 */
          8.33333333332248946124e-03,   /* 4  S2 */
          -1.98412698298579493134e-04,  /* 5  S3 */
          2.75573137070700676789e-06,   /* 6  S4 */
          -2.50507602534068634195e-08,  /* 7  S5 */
          1.58969099521155010221e-10,   /* 8  S6 */
          4.16666666666666019037e-02,   /* 9  C1 */
          -1.38888888888741095749e-03,  /* 10 C2 */
          2.48015872894767294178e-05,   /* 11 C3 */
          -2.75573143513906633035e-07,  /* 12 C4 */
          2.08757232129817482790e-09,   /* 13 C5 */
          -1.13596475577881948265e-11,  /* 14 C6 */
          0.0)

// log(x) = k ln2 + log(1+f), fdlibm __ieee754_log coefficients
TRI_TABLE(kLog, 10,
          6.93147180369123816490e-01,   /* 0 ln2 high */
          1.90821492927058770002e-10,   /* 1 ln2 low  */
          6.666666666666735130e-01,     /* 2 Lg1 */
          3.999999999940941908e-01,     /* 3 Lg2 */
          2.857142874366239149e-01,     /* 4 Lg3 */
          2.222219843214978396e-01,     /* 5 Lg4 */
          1.818357216161805012e-01,     /* 6 Lg5 */
          1.531383769920937332e-01,     /* 7 Lg6 */
          1.479819860511658591e-01,     /* 8 Lg7 */
          0.0)

// Hastings K / E polynomials (A&S 17.3.34, 17.3.36): K = (a0..a4)(m1) - (b0..b4)(m1) ln m1,
// E = 1 + (c1..c4)(m1) - (d1..d4)(m1) ln m1
TRI_TABLE(kEll, 18,
          1.38629436112, 0.09666344259, 0.03590092383, 0.03742563713, 0.01451196212,
          0.5, 0.12498593597, 0.06880248576, 0.03328355346, 0.00441787012,
          0.44325141463, 0.0626060122, 0.04757383546, 0.01736506451,
          0.2499836831, 0.09200180037, 0.04069697526, 0.00526449639)

// acos: fdlibm's rational approximation R(z) = z P(z) / Q(z) of (asin(x) - x) / x
TRI_TABLE(kAcos, 12,
          1.66666666666666657415e-01,   /* 0 pS0 */
          -3.25565818622400915405e-01,  /* 1 pS1 */
          2.01212532134862925881e-01,   /* 2 pS2 */
          -4.00555345006794114027e-02,  /* 3 pS3 */
          7.91534994289814532176e-04,   /* 4 pS4 */
          3.47933107596021167570e-05,   /* 5 pS5 */
          -2.40339491173441421878e+00,  /* 6 qS1 */
          2.02094576023350569471e+00,   /* 7 qS2 */
          -6.88283971605453293030e-01,  /* 8 qS3 */
          7.70381505559019352791e-02,   /* 9 qS4 */
          1.57079632679489655800e+00,   /* 10 pi/2 high */
          6.12323399573676603587e-17)   /* 11 pi/2 low  */

// acos(x) for |x| <= 1 (the limb-crossing angles of occult_quad; NaN -> NaN), ~1 ulp: one
// reciprocal and one square root from the hardware seeds instead of the CUDA library routine's
// 138 instructions with 26 embedded literals.
TRI_HD double acos_unit(double x) {
#if defined(__CUDA_ARCH__)
    const double* T = TRI_T(kAcos);
    const double ax = fabs(x);
    const bool big = ax >= 0.5;
    const double z = big ? (1.0 - ax) * 0.5 : x * x;
    const double pz = z * (T[0] + z * (T[1] + z * (T[2] + z * (T[3] + z * (T[4] + z * T[5])))));
    const double qz = 1.0 + z * (T[6] + z * (T[7] + z * (T[8] + z * T[9])));
    const double r = pz * fast_rcp(qz);
    if (!big) return T[10] - (x - (T[11] - x * r));
    const double sq = fast_sqrt(z);      // (|x| > 1 -> NaN, as acos)
    const double t = 2.0 * fma(sq, r, sq);
    return x > 0.0 ? t : 2.0 * T[10] - t + 2.0 * T[11];
#else
    return acos(x);
#endif
}

// sin and cos of an angle of a few radians at most (true anomalies): quadrant reduction with a
// two-term pi/2, then the two polynomials; ~1 ulp, no huge-argument path (|x| < 1e5 assumed).
TRI_HD void sincos_small(double x, double& s, double& c) {
    const double* T = TRI_T(kTrig);
    double q = rint(x * T[0]);
    double r = fma(-q, T[1], x);
    r = fma(-q, T[2], r);
    double z = r * r;
    double sp = T[4] + z * (T[5] + z * (T[6] + z * (T[7] + z * T[8])));
    sp = fma(r * z, fma(z, sp, T[3]), r);                    // r + r^3 (S1 + z (...))
    double cp = T[9] + z * (T[10] + z * (T[11] + z * (T[12] + z * (T[13] + z * T[14]))));
    cp = fma(z * z, cp, fma(-0.5, z, 1.0));                  // 1 - z/2 + z^2 (C1 + ...)
    int n = (int)q;
    double ss = (n & 1) ? cp : sp;
    double cc = (n & 1) ? sp : cp;
    s = (n & 2) ? -ss : ss;
    c = ((n + 1) & 2) ? -cc : cc;
}

// natural logarithm of a normal positive double (0 -> -inf, negative / NaN -> NaN)
TRI_HD double log_pos(double x) {
    if (!(x > 0.0)) return x == 0.0 ? -INFINITY : NAN;
    const double* T = TRI_T(kLog);
    int k0 = 0;
    if (x < 2.2250738585072014e-308) {   // subnormal: rescale by 2^54
        x *= 18014398509481984.0;
        k0 = -54;
    }
#if defined(__CUDA_ARCH__)
    int hx = __double2hiint(x);
    int lx = __double2loint(x);
#else
    union { double d; unsigned long long u; } cv;
    cv.d = x;
    int hx = (int)(cv.u >> 32);
    int lx = (int)(cv.u & 0xffffffffu);
#endif
    int k = (hx >> 20) - 1023 + k0;
    hx &= 0x000fffff;
    int i = (hx + 0x95f64) & 0x100000;        // mantissa above sqrt(2): halve it, bump k
    hx |= (i ^ 0x3ff00000);
    k += (i >> 20);
#if defined(__CUDA_ARCH__)
    double m = __hiloint2double(hx, lx);
#else
    cv.u = ((unsigned long long)(unsigned)hx << 32) | (unsigned)lx;
    double m = cv.d;
#endif
    double f = m - 1.0;
    double sq = f * fast_rcp(2.0 + f);
    double dk = (double)k;
    double z = sq * sq;
    double w = z * z;
    double t1 = w * (T[3] + w * (T[5] + w * T[7]));
    double t2 = z * (T[2] + w * (T[4] + w * (T[6] + w * T[8])));
    double R = t2 + t1;
    double hfsq = 0.5 * f * f;
    return dk * T[0] - ((hfsq - (sq * (hfsq + R) + dk * T[1])) - f);
}

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
    const double* prefix;  // [npts+1] running sum of w (flux-1)^2, prefix[0]=0
    const double* weight;  // [npts] 1/sigma_j^2 for per-point errors (tri_set_lightcurve_err),
                           // nullptr for the reference's scalar sigma (w == 1 in the sums, the
                           // division by sigma^2 is applied once at the end, likelihoods.py:486)
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
    double esw, ecw;   // e sin w, e cos w: 1 + e cos f = 1 + ecw cos(w+f) + esw sin(w+f)
    bool table_clamped;  // e beyond the table: last cell extrapolated (monotonicity not guaranteed)
};

// Mean anomaly at mid-transit (true anomaly pi/2 - w), from sin w and cos w:
// sin(pi/2 - w) = cos w, cos(pi/2 - w) = sin w.
TRI_HD double mean_anomaly_offset(double e, double sinw, double cosw) {
    double off = atan2(sqrt(1.0 - e * e) * cosw, e + sinw);
    off -= e * sin(off);
    return off;
}

TRI_HD void orbit_setup(Orbit& o, const OrbitTable& T, double k, double p, double a_rs,
                        double inc_rad, double e, double w) {
    o.k = k;
    o.e = e;
    o.n_rate = kTwoPi / p;
    sincos(w, &o.sinw, &o.cosw);
    double off = mean_anomaly_offset(e, o.sinw, o.cosw);
    o.ma_tr = off;
    o.c0 = 0.0 - off * p / kTwoPi;
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
    o.esw = e * o.sinw;
    o.ecw = e * o.cosw;
}

// The functions of the hot loop take the orbit / limb record as a template parameter: the
// kernel keeps ONE copy per warp in shared memory and reads it through a `volatile` reference,
// so that every field is fetched (LDS, warp broadcast) where it is used instead of living in
// 32 x 2 registers for the whole draw; each field is read once per evaluation.
// table part of the true anomaly: f(M) for M already reduced to [0, 2 pi)
template <class OrbitT>
TRI_HD double ta_from_ma(const OrbitT& o, const OrbitTable& T, double ma) {
    double x, s;
    if (ma < kPi) { x = ma; s = 1.0; } else { x = kTwoPi - ma; s = -1.0; }
    // the blend is continuous across cells, so x*inv_dm landing one cell off the oracle's
    // floor(x/dm) at a cell edge only changes the result by rounding
    int im = (int)floor(x * T.inv_dm);
    if (im > kTableNm - 2) im = kTableNm - 2;
    double am = (x - im * T.dm) * T.inv_dm;
    const double* r0 = o.row0;
    const double* r1 = o.row1;
    const double ae = o.ae;
#if defined(__CUDA_ARCH__)
    double t00 = __ldg(r0 + im), t01 = __ldg(r0 + im + 1);
    double t10 = __ldg(r1 + im), t11 = __ldg(r1 + im + 1);
#else
    double t00 = r0[im], t01 = r0[im + 1], t10 = r1[im], t11 = r1[im + 1];
#endif
    // bilinear blend as three lerps (6 FP64 ops; the oracle's four-product form differs from it
    // by rounding only)
    double c0 = fma(ae, t10 - t00, t00);
    double c1 = fma(ae, t11 - t01, t01);
    double d = fma(am, c1 - c0, c0);
    return fma(s, d, ma);
}

template <class OrbitT>
TRI_HD double mean_anomaly(const OrbitT& o, double t) {
    double x = (t - o.c0) * o.n_rate;
    double ma = fma(-kTwoPi, floor(x * kInvTwoPi), x);
    if (ma < 0.0) ma += kTwoPi;
    if (ma >= kTwoPi) ma -= kTwoPi;
    return ma;
}

template <class OrbitT>
TRI_HD double z_from_ta(const OrbitT& o, double ta) {
    double st, ct;
    sincos_small(ta, st, ct);
    double swt = o.sinw * ct + o.cosw * st;  // sin(w + f)
    double z = o.a1me2 * fast_rcp(1.0 + o.e * ct) * fast_sqrt(1.0 - swt * swt * o.sini2);
    return swt < 0.0 ? -z : z;
}

template <class OrbitT>
TRI_HD double z_at(const OrbitT& o, const OrbitTable& T, double t) {
    return z_from_ta(o, ta_from_ma(o, T, mean_anomaly(o, t)));
}

// ---- sub-exposures of one time stamp ------------------------------------------------------------
// The sub-exposures of a stamp lie within exptime/2 of its centre, so most of the orbit work of
// z_at() can be shared: the mean anomaly is the centre's plus n*dt (no range reduction: the
// caller checks once per stamp that the exposure stays inside one half-turn of the (e, M) table),
// and sin / cos of w + f follow from the centre's by angle addition with the SMALL true-anomaly
// difference d (two short Taylor polynomials, no quadrant logic).  The interpolated f(M) itself
// is evaluated per sub-exposure exactly as in ta_from_ma(): table error is part of the model.
// When d leaves the polynomials' range (long exposures on eccentric orbits) the base point
// moves to the current sub-exposure.
TRI_TABLE(kTaylor, 10,
          -1.66666666666666666667e-01,  /* 0 sin: -1/3!  */
          8.33333333333333333333e-03,   /* 1       1/5!  */
          -1.98412698412698412698e-04,  /* 2      -1/7!  */
          2.75573192239858906526e-06,   /* 3       1/9!  */
          -2.50521083854417187751e-08,  /* 4      -1/11! */
          -5.00000000000000000000e-01,  /* 5 cos: -1/2!  */
          4.16666666666666666667e-02,   /* 6       1/4!  */
          -1.38888888888888888889e-03,  /* 7      -1/6!  */
          2.48015873015873015873e-05,   /* 8       1/8!  */
          -2.75573192239858906526e-07)  /* 9      -1/10! */

constexpr double kStampMaxDelta = 0.125;   // |d| beyond which the base point is moved
                                           // (truncation: d^12/12! < 3e-20, d^13/13! < 3e-22)

struct StampOrbit {
    double ma_c;     // mean anomaly at the stamp centre, reduced to [0, 2 pi)
    double ta_b;     // true anomaly of the base point
    double S_b, C_b; // sin(w + ta_b), cos(w + ta_b)
    double sgn, xoff; // table argument x = xoff + sgn * M: (0, +1) on [0, pi), (2 pi, -1) on
                      // [pi, 2 pi) where the table is read mirrored; f = M + sgn * (f - M)(x)
};

// z from sin(w+f), cos(w+f)
template <class OrbitT>
TRI_HD double z_from_sc(const OrbitT& o, double S, double C) {
    double den = fma(o.ecw, C, fma(o.esw, S, 1.0));          // 1 + e cos f
    double z = o.a1me2 * fast_rcp(den) * fast_sqrt_clamped(1.0 - S * S * o.sini2);
    return S < 0.0 ? -z : z;
}

// Centre of a stamp: the generic evaluation (range reduction, table, full sincos); fills `so`
// and returns z.  `fast` tells whether every sub-exposure within +-half_ma of the centre's mean
// anomaly stays inside one half-turn, i.e. may use z_sub().
template <class OrbitT>
TRI_HD double stamp_centre(const OrbitT& o, const OrbitTable& T, double t, double half_ma,
                           StampOrbit& so, bool& fast) {
    const double ma = mean_anomaly(o, t);
    const double ta = ta_from_ma(o, T, ma);
    double st, ct;
    sincos_small(ta, st, ct);
    const double sw = o.sinw, cw = o.cosw;
    so.ma_c = ma;
    so.ta_b = ta;
    so.S_b = sw * ct + cw * st;
    so.C_b = cw * ct - sw * st;
    const bool upper = ma >= kPi;
    so.sgn = upper ? -1.0 : 1.0;
    so.xoff = upper ? kTwoPi : 0.0;
    const double lo = ma - half_ma, hi = ma + half_ma;
    fast = upper ? (lo > kPi && hi < kTwoPi) : (lo > 0.0 && hi < kPi);
    return z_from_sc(o, so.S_b, so.C_b);
}

// Sub-exposure at time offset `toff` from the centre of a stamp prepared by stamp_centre()
// (only when it reported fast).
template <class OrbitT>
TRI_HD double z_sub(const OrbitT& o, const OrbitTable& T, StampOrbit& so, double toff) {
    const double ma = fma(toff, o.n_rate, so.ma_c);
    // 0 < x < pi strictly (stamp_centre's check), so the cell index is at most kTableNm - 2
    // except when x inv_dm rounds up to kTableNm - 1 exactly: the weight of the next cell is
    // then 0 and the table carries one padding element behind its last row
    const double x = fma(so.sgn, ma, so.xoff);
    const double u = x * T.inv_dm;
    const double fl = floor(u);
    const double am = u - fl;
    const int im = (int)fl;
    const double* r0 = o.row0;
    const double* r1 = o.row1;
    const double ae = o.ae;
#if defined(__CUDA_ARCH__)
    double t00 = __ldg(r0 + im), t01 = __ldg(r0 + im + 1);
    double t10 = __ldg(r1 + im), t11 = __ldg(r1 + im + 1);
#else
    double t00 = r0[im], t01 = r0[im + 1], t10 = r1[im], t11 = r1[im + 1];
#endif
    const double c0 = fma(ae, t10 - t00, t00);
    const double c1 = fma(ae, t11 - t01, t01);
    const double d = fma(am, c1 - c0, c0);
    const double ta = fma(so.sgn, d, ma);
    const double dl = ta - so.ta_b;
    double S, C;
    if (fabs(dl) > kStampMaxDelta) {   // rare: move the base point here
        double st, ct;
        sincos_small(ta, st, ct);
        const double sw = o.sinw, cw = o.cosw;
        S = sw * ct + cw * st;
        C = cw * ct - sw * st;
        so.ta_b = ta;
        so.S_b = S;
        so.C_b = C;
    } else {
        const double* K = TRI_T(kTaylor);
        const double z2 = dl * dl;
        double sp = K[1] + z2 * (K[2] + z2 * (K[3] + z2 * K[4]));
        sp = fma(dl * z2, fma(z2, sp, K[0]), dl);                  // sin d
        double cp = K[6] + z2 * (K[7] + z2 * (K[8] + z2 * K[9]));
        cp = fma(z2, fma(z2, cp, K[5]), 1.0);                      // cos d
        S = fma(so.S_b, cp, so.C_b * sp);
        C = fma(so.C_b, cp, -(so.S_b * sp));
    }
    return z_from_sc(o, S, C);
}

// ---------------------------------------------------------------- elliptic integrals
// K and E from the complementary parameter m1 = 1 - q^2; they share its logarithm
// (Hastings, A&S 17.3.34 / 17.3.36).
TRI_HD void ellke_m1(double m1, double& Kk, double& Ek) {
    const double* T = TRI_T(kEll);
    double lg = log_pos(m1);
    double ek1 = T[0] + m1 * (T[1] + m1 * (T[2] + m1 * (T[3] + m1 * T[4])));
    double ek2 = T[5] + m1 * (T[6] + m1 * (T[7] + m1 * (T[8] + m1 * T[9])));
    double ee1 = 1.0 + m1 * (T[10] + m1 * (T[11] + m1 * (T[12] + m1 * T[13])));
    double ee2 = m1 * (T[14] + m1 * (T[15] + m1 * (T[16] + m1 * T[17])));
    Kk = ek1 - ek2 * lg;
    Ek = ee1 - ee2 * lg;
}

// Bulirsch (1965) third-kind integral, started from kc = sqrt(1 - q^2), p = sqrt(n + 1) and
// d = 1/p (the callers have closed forms for p and d, see occult_quad).  One reciprocal per
// sweep instead of two divisions.  NaN inputs return NaN, as in the oracle (a NaN's high word
// is above every threshold: one sweep).
TRI_HD double ellpicb(double kc, double p, double d) {
    // Sweeps until the oracle's test |1 - kc/g| <= 1e-8 holds: a function of the starting kc
    // alone (the (m0, kc) pair is the AGM of (1, kc), scaled), so it is read off the high word
    // of kc instead of being tested in every sweep (thresholds rounded up by one unit of the
    // high word: a sweep too many changes the value by < 7e-16 relative, a sweep too few
    // cannot happen).  kc < 2^-86 can only be kc == 0, where the oracle's loop never converges
    // and returns 0 after its cap.
    const unsigned h = hi_word(kc);
    if (h < 0x3a900000u) return 0.0;
    int it = 1 + (h < 0x3ff00000u) + (h < 0x3feffdafu) + (h < 0x3fee836eu) + (h < 0x3fe12ee9u)
           + (h < 0x3fb5b7c9u) + (h < 0x3f5d95f7u) + (h < 0x3eab5a91u) + (h < 0x3d4761d4u);
    double e = kc;
    double m0 = 1.0, c = 1.0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (;;) {
        double ip = fast_rcp(p);
        double f = c;
        c = fma(d, ip, c);
        double g = e * ip;
        d = 2.0 * fma(f, g, d);
        p = g + p;
        m0 = kc + m0;
        if (--it == 0) break;
        kc = 2.0 * fast_sqrt_pos(e);
        e = kc * m0;
    }
    return kHalfPi * fma(c, m0, d) * fast_rcp(m0 * (m0 + p));
}

// ---------------------------------------------------------------------- occultation
// Limb-darkening mix for one sample: flux = 1 - (c_le*le + c_ld*ld + u2*ed) * inv_omega
struct Limb {
    double c_le, c_ld, u2, inv_omega;
    double inv_k;   // 1/k of the occultor this mix is used with
};

TRI_HD void limb_setup(Limb& L, double u1, double u2, double k) {
    L.c_le = 1.0 - u1 - 2.0 * u2;
    L.c_ld = u1 + 2.0 * u2;
    L.u2 = u2;
    L.inv_omega = 1.0 / (1.0 - u1 / 3.0 - u2 / 6.0);
    L.inv_k = 1.0 / k;
}

// Relative flux at separation z (any sign), radius ratio k.  Mirrors oracle eval_quad(): same
// case tests, same formulas, with the algebra arranged for the GPU:
//   * K, E and Pi are evaluated once for whichever of the two general cases (limb-crossing III /
//     interior IV) applies, so divergent lanes share them;
//   * m1 = 1 - q^2 is formed directly (the oracle takes q = sqrt(..) and squares it again);
//   * the third-kind parameter n only enters through p = sqrt(n + 1), which is
//     (k+z)/|k-z| in case IV and 1/|k-z| in case III, and 1/x1 follows from the same
//     reciprocal: no square root or extra division is needed for them.
// `cls` receives the work class of SURVEY.md 8(d): 0 = trivial (out of transit, behind the
// star, total eclipse), 1 = interior (z <= 1 - k), 2 = limb-crossing.
template <class LimbT>
TRI_HD double occult_quad(double z, double k, const LimbT& L, int& cls) {
    cls = 0;
    if (fabs(z - k) < 1e-6) z += 1e-6;
    if (z > 1.0 + k || z < 0.0) return 1.0;
    if (k >= 1.0 && z <= k - 1.0) return 0.0;

    const double k2 = k * k, z2 = z * z;
    const double dkz = k - z, skz = k + z;
    const double x1 = dkz * dkz, x2 = skz * skz, x3 = k * k - z * z;
    double le = 0.0, ld = 0.0, ed = 0.0, kap0 = 0.0, kap1 = 0.0;

    const bool partial = (z >= fabs(1.0 - k) && z <= 1.0 + k);
    cls = partial ? 2 : 1;
    if (partial) {
        double iz = fast_rcp(z);
        kap1 = acos_unit(fmin((1.0 - k2 + z2) * 0.5 * iz, 1.0));
        kap0 = acos_unit(fmin((k2 + z2 - 1.0) * 0.5 * iz * L.inv_k, 1.0));
        double t = 1.0 + z2 - k2;
        le = (k2 * kap0 + kap1 - 0.5 * sqrt(fmax(4.0 * z2 - t * t, 0.0))) * kInvPi;
    }
    if (z <= 1.0 - k) le = k2;

    const bool edge = fabs(dkz) < 1e-4 * skz;
    const bool case3 = !edge && ((z > 0.5 + fabs(k - 0.5) && z < 1.0 + k)
                                 || (k > 0.5 && z > fabs(1.0 - k) * 1.0001 && z < k));
    const bool case4 = !edge && !case3 && (k <= 1.0 && z < (1.0 - k) * 1.0001);

    const bool edge_ke = edge && (k != 0.5);
    if (case3 || case4 || edge_ke) {
        const double om = 1.0 - x1;
        const double adk = fabs(dkz);
        double m1, p = 1.0, d = 1.0, rs = 1.0, ix1 = 0.0;
        if (case3) {
            // q^2 = (1-x1)/(4kz);  n + 1 = 1/x1
            rs = fast_rsqrt(k * z);                 // 1/sqrt(kz)
            m1 = 1.0 - om * 0.25 * rs * rs;
            p = fast_rcp(adk);
            d = adk;
            ix1 = p * p;                            // 1/x1
        } else if (case4) {
            // q^2 = (x2-x1)/(1-x1);  n + 1 = x2/x1
            rs = fast_rsqrt(om);                    // 1/sqrt(1-x1)
            m1 = 1.0 - (x2 - x1) * rs * rs;
            double i3 = fast_rcp(adk * skz);        // 1/|k^2 - z^2|
            p = x2 * i3;
            d = x1 * i3;
            ix1 = i3 * p;                           // 1/x1
        } else {
            // occultor's edge at the disc centre: q = 1/(2k) (z > 1/2) or 2k (z < 1/2)
            double q = (z > 0.5) ? 0.5 * L.inv_k : 2.0 * k;
            m1 = 1.0 - q * q;
        }
        double Kk, Ek;
        ellke_m1(m1, Kk, Ek);
        if (!edge_ke) {
            const double Pk = ellpicb(fast_sqrt(m1), p, d);
            // x3 keeps the oracle's k*k - z*z rounding (it cancels near z = k)
            const double spterm = 3.0 * x3 * ix1 * Pk;
            if (case3) {
                ld = 1.0 / 9.0 * kInvPi * rs
                   * (((1.0 - x2) * (2.0 * x2 + x1 - 3.0) - 3.0 * x3 * (x2 - 2.0)) * Kk
                      + 4.0 * k * z * (z2 + 7.0 * k2 - 4.0) * Ek - spterm);
                if (z < k) ld += 2.0 / 3.0;
            } else {
                ld = 2.0 / 9.0 * kInvPi * rs
                   * ((1.0 - 5.0 * z2 + k2 + x3 * x3) * Kk
                      + om * (z2 + 7.0 * k2 - 4.0) * Ek - spterm);
                if (z < k) ld += 2.0 / 3.0;
                if (fabs(k + z - 1.0) < 1e-4)
                    ld = 2.0 / 3.0 * kInvPi * acos(1.0 - 2.0 * k)
                       - 4.0 / 9.0 * kInvPi * sqrt(k * (1.0 - k)) * (3.0 + 2.0 * k - 8.0 * k2);
            }
        } else if (z > 0.5) {
            ld = 1.0 / 3.0 + 16.0 * k / 9.0 * kInvPi * (2.0 * k2 - 1.0) * Ek
               - (32.0 * (k2 * k2) - 20.0 * k2 + 3.0) / 9.0 * kInvPi * L.inv_k * Kk;
        } else {
            ld = 1.0 / 3.0 + 2.0 / 9.0 * kInvPi * (4.0 * (2.0 * k2 - 1.0) * Ek
               + (1.0 - 4.0 * k2) * Kk);
        }
        // eta: limb-crossing form when the occultor reaches beyond the disc centre side
        if (case3 || (edge_ke && z > 0.5))
            ed = 0.5 * kInvPi * (kap1 + k2 * (k2 + 2.0 * z2) * kap0
               - (1.0 + 5.0 * k2 + z2) * 0.25 * sqrt((1.0 - x1) * (x2 - 1.0)));
        else
            ed = k2 * 0.5 * (k2 + 2.0 * z2);
    } else if (edge) {   // k == 1/2 exactly
        ld = 1.0 / 3.0 - 4.0 * kInvPi / 9.0;
        ed = 3.0 / 32.0;
    }
    return 1.0 - (L.c_le * le + L.c_ld * ld + L.u2 * ed) * L.inv_omega;
}

template <class LimbT>
TRI_HD double occult_quad(double z, double k, const LimbT& L) {
    int cls;
    return occult_quad(z, k, L, cls);
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
    sincos_small(ta, st, ct);
    double swt = o.sinw * ct + o.cosw * st;
    double cwt = o.cosw * ct - o.sinw * st;
    return swt > 0.0 && fabs(cwt) <= cmax;
}

TRI_HD bool transit_window(const Orbit& o, const OrbitTable& T, double a_rs, double p,
                           const LightCurve& lc, Window& win) {
    if (o.table_clamped) return false;
    // z^2 = r^2 (1 - sin^2(w+f) sin^2 i) <= (1+k)^2 with r >= a(1-e) bounds the in-transit arc:
    // cos^2(w+f) <= (c0^2 - cos^2 i) / sin^2 i, c0 = (1+k)/r_min.  (Grazing and high-impact
    // draws get the short window their chord deserves; 1e-12 covers the cancellation.)
    double rmin = a_rs * (1.0 - o.e);
    double c0 = (1.0 + o.k) / rmin;
    double num = c0 * c0 - (1.0 - o.sini2);
    double cmax = sqrt(fmax(num, 0.0) / o.sini2 + 1e-12) * (1.0 + 1e-9);
    if (!(cmax < 0.95)) return false;  // wide arcs: not worth it / not safe
    // f(M) is monotone (bilinear blend of monotone rows); M is taken relative to mid-transit.
    // The centre must be inside the arc and the opposite point outside; each edge is then the
    // first offset known to be outside.
    double ma0 = o.ma_tr - kTwoPi * floor(o.ma_tr * kInvTwoPi);
    double edge[2];
#if defined(__CUDA_ARCH__)
    // The whole warp works on one draw: a 17-ary search, 16 lanes per side, replaces 2 x 28
    // serial bisection steps by 7 evaluations (resolution pi / 17^6 = 1.3e-7 rad).
    {
        const unsigned lane = threadIdx.x & 31u;
        const int side = (int)(lane >> 4);
        const int q = (int)(lane & 15u);
        const double sgn = side ? 1.0 : -1.0;
        double lo = 0.0, hi = kPi;  // offset from mid-transit; lo inside the arc, hi outside
#pragma unroll 1
        for (int it = -1; it < 6; ++it) {
            double mid = (it < 0) ? (lane == 0 ? 0.0 : kPi)
                                  : lo + (hi - lo) * ((double)(q + 1) * (1.0 / 17.0));
            double m = ma0 + sgn * mid;
            m -= kTwoPi * floor(m * kInvTwoPi);
            if (m >= kTwoPi) m -= kTwoPi;
            if (m < 0.0) m += kTwoPi;
            const bool inside = in_arc(o, ta_from_ma(o, T, m), cmax);
            const unsigned b = __ballot_sync(0xffffffffu, inside);
            if (it < 0) {   // lane 0 probed the centre, lane 1 the opposite point
                if ((b & 3u) != 1u) return false;
                continue;
            }
            // first lane of this side that is outside (16: none)
            const unsigned out = ~(b >> (side * 16)) & 0xffffu;
            const int c = out ? (__ffs(out) - 1) : 16;
            const double up = __shfl_sync(0xffffffffu, mid, side * 16 + min(c, 15));
            const double dn = __shfl_sync(0xffffffffu, mid, side * 16 + max(c - 1, 0));
            if (c < 16) hi = up;
            if (c > 0) lo = dn;
        }
        edge[0] = __shfl_sync(0xffffffffu, hi, 0);
        edge[1] = __shfl_sync(0xffffffffu, hi, 16);
    }
#else
    for (int side = 0; side < 2; ++side) {
        double sgn = side ? 1.0 : -1.0;
        double lo = 0.0, hi = kPi;  // offset from mid-transit; lo inside the arc, hi outside
        for (int it = -2; it < 26; ++it) {
            double mid = (it == -2) ? 0.0 : (it == -1) ? kPi : 0.5 * (lo + hi);
            double m = ma0 + sgn * mid;
            m -= kTwoPi * floor(m * kInvTwoPi);
            if (m >= kTwoPi) m -= kTwoPi;
            if (m < 0.0) m += kTwoPi;
            bool inside = in_arc(o, ta_from_ma(o, T, m), cmax);
            if (it == -2) { if (!inside) return false; }
            else if (it == -1) { if (inside) return false; }
            else if (inside) lo = mid;
            else hi = mid;
        }
        edge[side] = hi;  // first offset known to be outside
    }
#endif
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

// Bound on |z(t1) - z(t2)| / |t1 - t2| [stellar radii per day] for the interpolated orbit: the
// projected separation cannot change faster than the body moves, |dr/df| * df/dt, with
// |dr/df| <= a(1+e) sqrt(1 + e^2/(1-e^2)) on the ellipse and df/dt <= n (1+e')^2/(1-e'^2)^1.5
// where e' covers the next row of the (e, M) table.  Loose for eccentric orbits, always safe.
// A time stamp whose centre has |z| > 1 + k + speed * exptime/2 has every sub-exposure out of
// transit, so its model is exactly 1 and the sub-exposures need not be evaluated.
TRI_HD double max_projected_speed(const Orbit& o, double a_rs) {
    double e = o.e;
    double e1 = fmin(e + 0.004, 0.96);
    double om = 1.0 - e1 * e1;
    double dfdt = o.n_rate * (1.0 + e1) * (1.0 + e1) / (om * sqrt(om));
    double drdf = a_rs * (1.0 + e) * sqrt(1.0 + e * e / (1.0 - e * e));
    return 1.1 * dfdt * drdf;
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
