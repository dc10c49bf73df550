// TEST-ONLY: runs the __host__ __device__ model functions of
// triceratops_b200/csrc/tri_model.cuh serially on the CPU so that CPU tests can compare the
// kernel's arithmetic *structure* (transit window, merged case III/IV evaluation, reciprocal
// Bulirsch sweep, prefix-sum treatment of out-of-transit stamps) with the oracle without a GPU.
// This file is never linked into libtriceratops_b200.so and nothing in the package loads it.
#include <cmath>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <numeric>

#include "../../triceratops_b200/csrc/tri_model.cuh"

using namespace tri;

namespace {
std::vector<double> g_tae;
OrbitTable g_tab;

double ta_newton(double ma, double e) {
    double ea = ma, err = 0.05;
    int k = 0;
    while (std::fabs(err) > 1e-8 && k < 1000) {
        err = ea - e * std::sin(ea) - ma;
        ea = ea - err / (1.0 - e * std::cos(ea));
        k++;
    }
    double sta = std::sqrt(1.0 - e * e) * std::sin(ea) / (1.0 - e * std::cos(ea));
    double cta = (std::cos(ea) - e) / (1.0 - e * std::cos(ea));
    return std::atan2(sta, cta);
}

void ensure_table() {
    if (!g_tae.empty()) return;
    g_tae.assign((size_t)kTableNe * kTableNm + 1, 0.0);   // (+1 padding, as tri_init)
    double de = kTableMaxE / (kTableNe - 1), dm = kPi / (kTableNm - 1);
    for (int i = 0; i < kTableNe; i++) {
        double e = (i == kTableNe - 1) ? kTableMaxE : i * de;
        for (int j = 0; j < kTableNm; j++) {
            double m = (j == kTableNm - 1) ? kPi : j * dm;
            g_tae[(size_t)i * kTableNm + j] = ta_newton(m, e) - m;
        }
    }
    g_tab.tae = g_tae.data();
    g_tab.de = (1 * de) - 0.0;
    g_tab.dm = (1 * dm) - 0.0;
    g_tab.inv_dm = 1.0 / g_tab.dm;
}
}  // namespace

extern "C" {

double hc_occult_quad(double z, double k, double u1, double u2) {
    Limb L;
    limb_setup(L, u1, u2, k);
    return occult_quad(z, k, L);
}

double hc_z(double t, double p, double a, double inc, double e, double w) {
    ensure_table();
    Orbit o;
    orbit_setup(o, g_tab, 0.1, p, a, inc, e, w);
    return z_at(o, g_tab, t);
}

// mirrors lnl_kernel for one sample; returns +0.5 chi^2 (or +inf on the depth cut)
// stats[0] += stamps evaluated, stats[1] += 1 if a window was used
void hc_lnl(int eb, int64_t npts, const double* time_sorted, const double* flux_sorted,
            double sigma, double exptime, int nsamples, int64_t n, const double* body,
            const double* ebfr, const double* P, const double* inc_deg, const double* a_cm,
            const double* R_s, const double* u1, const double* u2, const double* ecc,
            const double* argp, const double* cfr_, int is_host, int twin, int use_window,
            double* out, int64_t* stats) {
    ensure_table();
    std::vector<double> pre(npts + 1);
    long double run = 0;
    pre[0] = 0;
    for (int64_t j = 0; j < npts; j++) {
        long double d = (long double)flux_sorted[j] - 1.0L;
        run += (long double)((double)d * (double)d);
        pre[j + 1] = (double)run;
    }
    LightCurve lc{time_sorted, flux_sorted, pre.data(), nullptr, (int)npts, nsamples, sigma, exptime,
                  time_sorted[0], time_sorted[npts - 1]};
    const double inv_ns = 1.0 / nsamples;
    for (int64_t i = 0; i < n; i++) {
        const double rhost = R_s[i];
        const double a_rs = a_cm[i] / (rhost * kRsun);
        const double inc = inc_deg[i] * (kPi / 180.0);
        const double w_rad = (90.0 - argp[i]) * (kPi / 180.0);
        const double cfr = cfr_[i];
        const double F_comp = cfr / (1.0 - cfr);
        Limb L;
        Dilution D;
        double k;
        bool cut = false;
        if (!eb) {
            k = body[i] * kRearth / (rhost * kRsun);
            limb_setup(L, u1[i], u2[i], k);
            D.two_stage = false;
            D.d1 = 0;
            D.d2 = is_host ? 1.0 / F_comp : F_comp / 1.0;
        } else {
            const double reb = body[i], fr = ebfr[i];
            const double F_EB = fr / (1.0 - fr);
            k = reb / rhost;
            if ((k - 1.0) < 1e-6) k *= 0.999;
            double ks = rhost / reb;
            if ((ks - 1.0) < 1e-6) ks *= 0.999;
            const double ws = (90.0 - argp[i] + 180.0) * (kPi / 180.0);
            Orbit os;
            orbit_setup(os, g_tab, ks, P[i], a_rs, inc, ecc[i], ws);
            Limb Ls;
            limb_setup(Ls, u1[i], u2[i], ks);
            double sec = INFINITY;
            for (int lane = 0; lane < 25; lane++) {
                double ts = (lane == 24) ? 0.05 : -0.05 + lane * ((0.05 - -0.05) / 24.0);
                double z = z_at(os, g_tab, ts);
                double m = (z > 1.0 + ks) ? 1.0 : occult_quad(z, ks, Ls);
                sec = std::fmin(sec, m);
            }
            limb_setup(L, u1[i], u2[i], k);
            D.two_stage = true;
            if (is_host) {
                D.d1 = F_EB / F_comp;
                sec = (sec + F_comp / F_EB) / (1.0 + F_comp / F_EB);
                D.d2 = 1.0 / (F_comp + F_EB);
            } else {
                D.d1 = F_EB / 1.0;
                sec = (sec + 1.0 / F_EB) / (1.0 + 1.0 / F_EB);
                D.d2 = F_comp / (1.0 + F_EB);
            }
            double sd = 1.0 - (sec + D.d2) / (1.0 + D.d2);
            cut = !twin && !(sd < 1.5 * sigma);
        }
        if (cut) { out[i] = INFINITY; continue; }
        Orbit o;
        orbit_setup(o, g_tab, k, P[i], a_rs, inc, ecc[i], w_rad);
        int jlo = 0, jhi = (int)npts;
        Window win;
        if (use_window && transit_window(o, g_tab, a_rs, P[i], lc, win)) {
            double half = 0.5 * exptime;
            jlo = lower_bound(lc.time, lc.npts, win.t_lo - half);
            jhi = lower_bound(lc.time, lc.npts, win.t_hi + half);
            if (jhi < jlo) jhi = jlo;
            if (stats) stats[1]++;
        }
        double chi = 0.0;
        const bool probe = nsamples > 1 && !o.table_clamped;
        const double skip_beyond =
            1.0 + k + max_projected_speed(o, a_rs) * (0.5 * exptime) + 1e-9;
        const double half_ma = o.table_clamped
            ? 1e30 : o.n_rate * (0.5 * exptime) * (1.0 + 1e-9) + 1e-12;
        for (int j = jlo; j < jhi; j++) {
            double t = lc.time[j], acc = 0.0;
            // as lnl_kernel: generic evaluation at the stamp centre, sub-exposures expanded
            // around it (z_sub) unless the exposure straddles a half-turn of the table
            StampOrbit so;
            bool fast;
            const double zc = stamp_centre(o, g_tab, t, half_ma, so, fast);
            if (probe && std::fabs(zc) > skip_beyond) {
                acc = (double)nsamples;
                if (stats) stats[2]++;
            } else {
                for (int is = 1; is <= nsamples; ++is) {
                    double toff = exptime * ((is - 0.5) * inv_ns - 0.5);
                    double z = fast ? z_sub(o, g_tab, so, toff) : z_at(o, g_tab, t + toff);
                    acc += (z > 1.0 + k) ? 1.0 : occult_quad(z, k, L);
                }
            }
            double m = dilute(D, acc / nsamples);
            double r = lc.flux[j] - m;
            chi = std::fma(r, r, chi);
        }
        chi += (pre[jlo] - pre[0]) + (pre[npts] - pre[jhi]);
        out[i] = 0.5 * (chi / (sigma * sigma));
        if (stats) stats[0] += jhi - jlo;
    }
}
}
