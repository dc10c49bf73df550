/*
 * ORACLE (test infrastructure, not product code) -- plain-C restatement of the reference's
 * per-sample likelihood seam, used as the bit-stable CPU comparator for the CUDA kernels and
 * as the timed CPU baseline (bench.py cpu_baseline / --impl reference, kind "port").
 *
 * PARITY UNPINNED: the transit-model arithmetic belongs to pytransit==2.2 (reference
 * setup.py:25), which is absent from /root/reference and from this image; see the header of
 * oracle/quadmodel.py, whose semantics this file restates one-to-one in C.
 *
 * What it follows in the reference (paths relative to /root/reference/triceratops/):
 *   tro_lnl_tp      likelihoods.py:443-487 (lnL_TP_p) over :302-358 (simulate_TP_transit_p)
 *   tro_lnl_eb      likelihoods.py:490-539 (lnL_EB_p) and :542-587 (lnL_EB_twin_p) over
 *                   :361-439 (simulate_EB_transit_p): 0.999 rule :406/:418, 25-point secondary
 *                   :417-424, two-stage dilution :427-438, secdepth cut :535-538
 *   tro_log_mean_exp  _numerics.py:12-51
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int tro_num_threads(void);

#define TRO_NE 256
#define TRO_NM 512
#define TRO_MAX_E 0.95

static const double PI = 3.14159265358979323846;
#define HALF_PI (0.5 * PI)
#define TWO_PI (2.0 * PI)
#define INV_PI (1.0 / PI)

/* astropy >= 4.0 constants in cgs (CODATA 2018 / IAU 2015), as oracle/shim/astropy/constants.py */
static const double C_RSUN = 69570000000.0;
static const double C_REARTH = 637810000.0;

static double g_es[TRO_NE], g_ms[TRO_NM], g_tae[TRO_NE * TRO_NM];
static int g_table_ready = 0;

/* ------------------------------------------------------------------ orbit */
static double ta_newton(double ma, double e) {
    double ea = ma, err = 0.05;
    int k = 0;
    while (fabs(err) > 1e-8 && k < 1000) {
        err = ea - e * sin(ea) - ma;
        ea = ea - err / (1.0 - e * cos(ea));
        k++;
    }
    double sta = sqrt(1.0 - e * e) * sin(ea) / (1.0 - e * cos(ea));
    double cta = (cos(ea) - e) / (1.0 - e * cos(ea));
    return atan2(sta, cta);
}

/* numpy.linspace(start, stop, n): start + i*step with step=(stop-start)/(n-1), last = stop */
static void linspace(double start, double stop, int n, double* out) {
    double step = (stop - start) / (double)(n - 1);
    for (int i = 0; i < n; i++) out[i] = start + (double)i * step;
    out[n - 1] = stop;
}

void tro_make_table(double* es, double* ms, double* tae) {
    if (!g_table_ready) {
        linspace(0.0, TRO_MAX_E, TRO_NE, g_es);
        linspace(0.0, PI, TRO_NM, g_ms);
        for (int i = 0; i < TRO_NE; i++)
            for (int j = 0; j < TRO_NM; j++)
                g_tae[i * TRO_NM + j] = ta_newton(g_ms[j], g_es[i]) - g_ms[j];
        g_table_ready = 1;
    }
    if (es) memcpy(es, g_es, sizeof g_es);
    if (ms) memcpy(ms, g_ms, sizeof g_ms);
    if (tae) memcpy(tae, g_tae, sizeof g_tae);
}

static double mean_anomaly_offset(double e, double w) {
    double off = atan2(sqrt(1.0 - e * e) * sin(HALF_PI - w), e + cos(HALF_PI - w));
    off -= e * sin(off);
    return off;
}

/* Python float modulo (result has the sign of the divisor) */
static double pymod(double x, double y) {
    double r = fmod(x, y);
    if (r != 0.0 && ((r < 0.0) != (y < 0.0))) r += y;
    return r;
}

static double z_ip(double t, double t0, double p, double a, double inc, double e, double w,
                   double off) {
    const double de = g_es[1] - g_es[0];
    const double dm = g_ms[1] - g_ms[0];
    int ie = (int)floor(e / de);
    if (ie > TRO_NE - 2) ie = TRO_NE - 2;
    double ae = (e - de * ie) / de;

    double ma = pymod(TWO_PI * (t - (t0 - off * p / TWO_PI)) / p, TWO_PI);
    double x, s;
    if (ma < PI) { x = ma; s = 1.0; } else { x = TWO_PI - ma; s = -1.0; }
    int im = (int)floor(x / dm);
    if (im > TRO_NM - 2) im = TRO_NM - 2;
    double am = (x - im * dm) / dm;
    const double* r0 = g_tae + ie * TRO_NM;
    const double* r1 = r0 + TRO_NM;
    double d = r0[im] * (1.0 - ae) * (1.0 - am) + r1[im] * ae * (1.0 - am)
             + r0[im + 1] * (1.0 - ae) * am + r1[im + 1] * ae * am;
    double ta = ma + s * d;

    double swt = sin(w + ta);
    double si = sin(inc);
    double z = a * (1.0 - e * e) / (1.0 + e * cos(ta)) * sqrt(1.0 - swt * swt * si * si);
    return swt < 0.0 ? -z : z;
}

/* ------------------------------------------------------- elliptic integrals */
static double ellk(double k) {
    double m1 = 1.0 - k * k;
    double ek1 = 1.38629436112 + m1 * (0.09666344259 + m1 * (0.03590092383
               + m1 * (0.03742563713 + m1 * 0.01451196212)));
    double ek2 = (0.5 + m1 * (0.12498593597 + m1 * (0.06880248576
               + m1 * (0.03328355346 + m1 * 0.00441787012)))) * log(m1);
    return ek1 - ek2;
}

static double ellec(double k) {
    double m1 = 1.0 - k * k;
    double ee1 = 1.0 + m1 * (0.44325141463 + m1 * (0.0626060122
               + m1 * (0.04757383546 + m1 * 0.01736506451)));
    double ee2 = m1 * (0.2499836831 + m1 * (0.09200180037 + m1 * (0.04069697526
               + m1 * 0.00526449639))) * log(1.0 / m1);
    return ee1 + ee2;
}

static double ellpicb(double n, double k) {
    double kc = sqrt(1.0 - k * k), e = kc, p = sqrt(n + 1.0), m0 = 1.0, c = 1.0, d = 1.0 / p;
    for (int it = 0; it < 1000; it++) {
        double f = c;
        c = d / p + c;
        double g = e / p;
        d = 2.0 * (f * g + d);
        p = g + p;
        g = m0;
        m0 = kc + m0;
        if (fabs(1.0 - kc / g) > 1e-8) {
            kc = 2.0 * sqrt(e);
            e = kc * m0;
        } else {
            return HALF_PI * (c * m0 + d) / (m0 * (m0 + p));
        }
    }
    return 0.0;
}

/* --------------------------------------------------------------- occultation */
/* class codes returned through *cls: 0 unocculted/total, 1 interior (case IV), 2 limb / edge */
static double eval_quad(double z, double k, double u1, double u2, int* cls) {
    *cls = 0;
    if (fabs(z - k) < 1e-6) z += 1e-6;
    if (z > 1.0 + k || z < 0.0) return 1.0;
    if (k >= 1.0 && z <= k - 1.0) return 0.0;

    double omega = 1.0 - u1 / 3.0 - u2 / 6.0;
    double k2 = k * k, z2 = z * z;
    double x1 = (k - z) * (k - z), x2 = (k + z) * (k + z), x3 = k * k - z * z;
    double le = 0.0, ld = 0.0, ed = 0.0, kap0 = 0.0, kap1 = 0.0;

    if (z >= fabs(1.0 - k) && z <= 1.0 + k) {
        kap1 = acos(fmin((1.0 - k2 + z2) / 2.0 / z, 1.0));
        kap0 = acos(fmin((k2 + z2 - 1.0) / 2.0 / k / z, 1.0));
        le = k2 * kap0 + kap1;
        double t = 1.0 + z2 - k2;
        le = (le - 0.5 * sqrt(fmax(4.0 * z2 - t * t, 0.0))) * INV_PI;
    }
    if (z <= 1.0 - k) le = k2;

    if (fabs(z - k) < 1e-4 * (z + k)) {
        *cls = 2;
        if (k == 0.5) {
            ld = 1.0 / 3.0 - 4.0 * INV_PI / 9.0;
            ed = 3.0 / 32.0;
        } else if (z > 0.5) {
            double q = 0.5 / k, Kk = ellk(q), Ek = ellec(q);
            ld = 1.0 / 3.0 + 16.0 * k / 9.0 * INV_PI * (2.0 * k2 - 1.0) * Ek
               - (32.0 * (k2 * k2) - 20.0 * k2 + 3.0) / 9.0 * INV_PI / k * Kk;
            ed = 1.0 / 2.0 * INV_PI * (kap1 + k2 * (k2 + 2.0 * z2) * kap0
               - (1.0 + 5.0 * k2 + z2) / 4.0 * sqrt((1.0 - x1) * (x2 - 1.0)));
        } else {
            double q = 2.0 * k, Kk = ellk(q), Ek = ellec(q);
            ld = 1.0 / 3.0 + 2.0 / 9.0 * INV_PI * (4.0 * (2.0 * k2 - 1.0) * Ek
               + (1.0 - 4.0 * k2) * Kk);
            ed = k2 / 2.0 * (k2 + 2.0 * z2);
        }
    } else if ((z > 0.5 + fabs(k - 0.5) && z < 1.0 + k)
               || (k > 0.5 && z > fabs(1.0 - k) * 1.0001 && z < k)) {
        *cls = 2;
        double q = sqrt((1.0 - x1) / 4.0 / z / k), Kk = ellk(q), Ek = ellec(q);
        double n = 1.0 / x1 - 1.0;
        double Pk = ellpicb(n, q);
        ld = 1.0 / 9.0 * INV_PI / sqrt(k * z)
           * (((1.0 - x2) * (2.0 * x2 + x1 - 3.0) - 3.0 * x3 * (x2 - 2.0)) * Kk
              + 4.0 * k * z * (z2 + 7.0 * k2 - 4.0) * Ek - 3.0 * x3 / x1 * Pk);
        if (z < k) ld += 2.0 / 3.0;
        ed = 1.0 / 2.0 * INV_PI * (kap1 + k2 * (k2 + 2.0 * z2) * kap0
           - (1.0 + 5.0 * k2 + z2) / 4.0 * sqrt((1.0 - x1) * (x2 - 1.0)));
    } else if (k <= 1.0 && z < (1.0 - k) * 1.0001) {
        *cls = 1;
        double q = sqrt((x2 - x1) / (1.0 - x1)), Kk = ellk(q), Ek = ellec(q);
        double n = x2 / x1 - 1.0;
        double Pk = ellpicb(n, q);
        ld = 2.0 / 9.0 * INV_PI / sqrt(1.0 - x1)
           * ((1.0 - 5.0 * z2 + k2 + x3 * x3) * Kk
              + (1.0 - x1) * (z2 + 7.0 * k2 - 4.0) * Ek - 3.0 * x3 / x1 * Pk);
        if (z < k) ld += 2.0 / 3.0;
        if (fabs(k + z - 1.0) < 1e-4)
            ld = 2.0 / 3.0 * INV_PI * acos(1.0 - 2.0 * k)
               - 4.0 / 9.0 * INV_PI * sqrt(k * (1.0 - k)) * (3.0 + 2.0 * k - 8.0 * k2);
        ed = k2 / 2.0 * (k2 + 2.0 * z2);
    }
    return 1.0 - ((1.0 - u1 - 2.0 * u2) * le + (u1 + 2.0 * u2) * ld + u2 * ed) / omega;
}

double tro_eval_quad(double z, double k, double u1, double u2) {
    int c;
    return eval_quad(z, k, u1, u2, &c);
}

double tro_z(double t, double p, double a, double inc, double e, double w) {
    tro_make_table(0, 0, 0);
    return z_ip(t, 0.0, p, a, inc, e, w, mean_anomaly_offset(e, w));
}

/* supersampled model flux of one sample at one time; counts[0..2] += class tallies */
static double model_point(double t, double k, double p, double a, double inc, double e, double w,
                          double off, double u1, double u2, double exptime, int ns,
                          int64_t* counts) {
    double acc = 0.0;
    for (int is = 1; is <= ns; is++) {
        double toff = exptime * ((is - 0.5) / ns - 0.5);
        double z = z_ip(t + toff, 0.0, p, a, inc, e, w, off);
        int cls = 0;
        if (z > 1.0 + k) acc += 1.0;
        else acc += eval_quad(z, k, u1, u2, &cls);
        if (counts) counts[cls]++;
    }
    return acc / ns;
}

/* model light curve of ONE sample (undiluted), for tests */
void tro_model(int64_t npts, const double* time, double k, double p, double a_rs, double inc_rad,
               double e, double w_rad, double u1, double u2, double exptime, int nsamples,
               double* out) {
    tro_make_table(0, 0, 0);
    double off = mean_anomaly_offset(e, w_rad);
    for (int64_t j = 0; j < npts; j++)
        out[j] = model_point(time[j], k, p, a_rs, inc_rad, e, w_rad, off, u1, u2, exptime,
                             nsamples, 0);
}

/* --------------------------------------------------------------- likelihoods */
/* lnL_TP_p: returns +0.5*chi2 per sample.  counts (may be NULL): int64[3] model-point classes */
void tro_lnl_tp(int64_t npts, const double* time, const double* flux, double sigma,
                double exptime, int nsamples, int64_t n, const double* R_p, const double* P_orb,
                const double* inc_deg, const double* a_cm, const double* R_s, const double* u1,
                const double* u2, const double* ecc, const double* argp_deg, const double* cfr,
                int companion_is_host, double* out, int64_t* counts) {
    tro_make_table(0, 0, 0);
    int64_t c0 = 0, c1 = 0, c2 = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : c0, c1, c2) num_threads(tro_num_threads())
    for (int64_t i = 0; i < n; i++) {
        int64_t cnt[3] = {0, 0, 0};
        double F_comp = cfr[i] / (1 - cfr[i]);
        double k = R_p[i] * C_REARTH / (R_s[i] * C_RSUN);
        double a = a_cm[i] / (R_s[i] * C_RSUN);
        double inc = inc_deg[i] * (PI / 180.);
        double w = (90 - argp_deg[i]) * (PI / 180.);
        double off = mean_anomaly_offset(ecc[i], w);
        double F_dilute = companion_is_host ? 1.0 / F_comp : F_comp / 1.0;
        double acc = 0.0;
        for (int64_t j = 0; j < npts; j++) {
            double m = model_point(time[j], k, P_orb[i], a, inc, ecc[i], w, off, u1[i], u2[i],
                                   exptime, nsamples, counts ? cnt : 0);
            m = (m + F_dilute) / (1 + F_dilute);
            double r = flux[j] - m;
            acc += r * r / (sigma * sigma);
        }
        out[i] = 0.5 * acc;
        c0 += cnt[0]; c1 += cnt[1]; c2 += cnt[2];
    }
    if (counts) { counts[0] += c0; counts[1] += c1; counts[2] += c2; }
}

/* lnL_EB_p (twin=0, +inf where secdepth >= 1.5 sigma) / lnL_EB_twin_p (twin=1).
 * secdepth_out (may be NULL) receives the per-sample secondary depth.
 * tro_lnl_eb_rule with scalar_rule = 1 is the scalar lnL_EB / lnL_EB_twin of the reference's
 * parallel=False loops: |k - 1| < 1e-6 and secondary k = 1/k (likelihoods.py:121-123, :137). */
void tro_lnl_eb_rule(int64_t npts, const double* time, const double* flux, double sigma,
                     double exptime, int nsamples, int64_t n, const double* R_EB,
                     const double* EB_fluxratio, const double* P_orb, const double* inc_deg,
                     const double* a_cm, const double* R_s, const double* u1, const double* u2,
                     const double* ecc, const double* argp_deg, const double* cfr,
                     int companion_is_host, int twin, int scalar_rule, double* out,
                     double* secdepth_out, int64_t* counts);

void tro_lnl_eb(int64_t npts, const double* time, const double* flux, double sigma,
                double exptime, int nsamples, int64_t n, const double* R_EB,
                const double* EB_fluxratio, const double* P_orb, const double* inc_deg,
                const double* a_cm, const double* R_s, const double* u1, const double* u2,
                const double* ecc, const double* argp_deg, const double* cfr,
                int companion_is_host, int twin, double* out, double* secdepth_out,
                int64_t* counts) {
    tro_lnl_eb_rule(npts, time, flux, sigma, exptime, nsamples, n, R_EB, EB_fluxratio, P_orb,
                    inc_deg, a_cm, R_s, u1, u2, ecc, argp_deg, cfr, companion_is_host, twin, 0,
                    out, secdepth_out, counts);
}

void tro_lnl_eb_rule(int64_t npts, const double* time, const double* flux, double sigma,
                double exptime, int nsamples, int64_t n, const double* R_EB,
                const double* EB_fluxratio, const double* P_orb, const double* inc_deg,
                const double* a_cm, const double* R_s, const double* u1, const double* u2,
                const double* ecc, const double* argp_deg, const double* cfr,
                int companion_is_host, int twin, int scalar_rule, double* out,
                double* secdepth_out, int64_t* counts) {
    tro_make_table(0, 0, 0);
    double tsec[25];
    linspace(-0.05, 0.05, 25, tsec);
    int64_t c0 = 0, c1 = 0, c2 = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : c0, c1, c2) num_threads(tro_num_threads())
    for (int64_t i = 0; i < n; i++) {
        int64_t cnt[3] = {0, 0, 0};
        double F_target = 1;
        double F_comp = cfr[i] / (1 - cfr[i]);
        double F_EB = EB_fluxratio[i] / (1 - EB_fluxratio[i]);
        double k = R_EB[i] / R_s[i];
        if (scalar_rule ? (fabs(k - 1.0) < 1e-6) : ((k - 1.0) < 1e-6)) k *= 0.999;
        double a = a_cm[i] / (R_s[i] * C_RSUN);
        double inc = inc_deg[i] * (PI / 180.);
        double w = (90 - argp_deg[i]) * (PI / 180.);
        double off = mean_anomaly_offset(ecc[i], w);
        /* secondary eclipse: roles swapped, 25 points, no supersampling */
        double ks = R_s[i] / R_EB[i];
        if (scalar_rule) ks = 1 / k;
        else if ((ks - 1.0) < 1e-6) ks *= 0.999;
        double ws = (90 - argp_deg[i] + 180) * (PI / 180.);
        double offs = mean_anomaly_offset(ecc[i], ws);
        double sec = INFINITY;
        for (int j = 0; j < 25; j++) {
            double m = model_point(tsec[j], ks, P_orb[i], a, inc, ecc[i], ws, offs, u1[i], u2[i],
                                   0.0, 1, 0);
            if (m < sec) sec = m;
        }
        double d1, F_dilute, secdepth;
        if (companion_is_host) {
            d1 = F_EB / F_comp;
            sec = (sec + F_comp / F_EB) / (1 + F_comp / F_EB);
            F_dilute = F_target / (F_comp + F_EB);
        } else {
            d1 = F_EB / F_target;
            sec = (sec + F_target / F_EB) / (1 + F_target / F_EB);
            F_dilute = F_comp / (F_target + F_EB);
        }
        secdepth = 1 - (sec + F_dilute) / (1 + F_dilute);
        if (secdepth_out) secdepth_out[i] = secdepth;
        if (!twin && !(secdepth < 1.5 * sigma)) {
            out[i] = INFINITY;
            continue;
        }
        double acc = 0.0;
        for (int64_t j = 0; j < npts; j++) {
            double m = model_point(time[j], k, P_orb[i], a, inc, ecc[i], w, off, u1[i], u2[i],
                                   exptime, nsamples, counts ? cnt : 0);
            m = (m + d1) / (1 + d1);
            m = (m + F_dilute) / (1 + F_dilute);
            double r = flux[j] - m;
            acc += r * r / (sigma * sigma);
        }
        out[i] = 0.5 * acc;
        c0 += cnt[0]; c1 += cnt[1]; c2 += cnt[2];
    }
    if (counts) { counts[0] += c0; counts[1] += c1; counts[2] += c2; }
}

/* _log_mean_exp (_numerics.py:12-51): NaN/-inf weigh zero but count in N; any +inf -> +inf */
double tro_log_mean_exp(const double* logw, int64_t n) {
    double m = -INFINITY;
    int any = 0;
    for (int64_t i = 0; i < n; i++) {
        if (isinf(logw[i]) && logw[i] > 0) return INFINITY;
        if (isfinite(logw[i])) { any = 1; if (logw[i] > m) m = logw[i]; }
    }
    if (!any) return -INFINITY;
    double s = 0.0;
    for (int64_t i = 0; i < n; i++)
        if (isfinite(logw[i])) s += exp(logw[i] - m);
    return m + log(s) - log((double)n);
}

/* Threads of the OpenMP loops over draws: 0 = the process-wide OpenMP setting.  bench.py's CPU
 * arms set it to the host's core count (torchrun exports OMP_NUM_THREADS=1 for every rank). */
static int g_threads = 0;

void tro_set_num_threads(int n) { g_threads = n > 0 ? n : 0; }

int tro_num_threads(void) {
    int n = 1;
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    n = g_threads > 0 ? g_threads : omp_get_max_threads();
#endif
    return n;
}
