// sm_100a kernels of the marginal-likelihood path.
//
//   geometry_tp/eb_kernel  one thread per prior draw: Kepler's-law semi-major axis, transit
//                          probability, collision and inclination masks (reference
//                          marginal_likelihoods.py:107-123 for TP-type, :254-299 for EB-type),
//                          and compaction of the surviving draws into a work list.
//   lnl_kernel             persistent warps; one warp per surviving draw.  The lanes sweep the
//                          time stamps inside the draw's transit window, each lane averaging its
//                          own sub-exposures; a shuffle tree reduces chi^2.  Points outside the
//                          window have model == 1 exactly and are added from a prefix sum of
//                          (flux-1)^2.  EB-type draws first evaluate the 25-point secondary
//                          eclipse (one point per lane) and its depth cut
//                          (likelihoods.py:417-438, :535-538).
//   lse_partial/final      log-mean-exp of lnL + lnprior as per-thread running (max, scaled sum)
//                          pairs merged per block and then across blocks (_numerics.py:12-51).
//
// FP64 CUDA-core work throughout: the path is not a contraction, so no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tri_model.cuh"

namespace tri {

// pointer + stride (0 = one value broadcast to every sample, 1 = per-sample array)
struct Col {
    const double* p;
    int64_t stride;
    __device__ __forceinline__ double at(int64_t i) const { return __ldg(p + i * stride); }
};

// ---- geometry ------------------------------------------------------------------------------
struct GeomTp {
    int64_t N;
    Col rp, P, inc, ecc, argp, mtot, rhost;
    const uint8_t* extra_mask;  // optional AND term (qs_comp != 0, logg/Teff cuts)
    double* a_out;              // [N] semi-major axis [cm] of surviving draws
    double* lnl_out;            // [N] pre-filled with -inf for the rejected draws
    uint8_t* mask_out;          // optional [N]
    int64_t* items;             // work list (sample index << 1), capacity N: long items are
                                // appended from the front, short ones from the back
    unsigned long long* n_items;  // [0] items at the front, [3] items at the back
};

struct GeomEb {
    int64_t N;
    Col reb, q, P, inc, ecc, argp, mtot, rhost;
    const uint8_t* extra_mask;
    double* a_out;      // a (q < 0.95) or a_twin (q >= 0.95) of surviving draws
    double* p_out;      // P or 2P
    double* lnl_out;    // [N] EB branch, -inf default
    double* lnl_twin_out;  // [N] twin branch, -inf default
    uint8_t* mask_out;       // optional [N]: mask of the EB branch
    uint8_t* mask_twin_out;  // optional [N]: mask of the twin branch
    int64_t* items;     // (sample index << 1) | twin; front / back as in GeomTp
    unsigned long long* n_items;  // [0] items at the front, [1] twins, [3] items at the back
};

__device__ __forceinline__ double neg_inf() { return -INFINITY; }

// a = ((G*M*Msun)/(4*pi**2)*(P*86400)**2)**(1/3)        marginal_likelihoods.py:75
__device__ __forceinline__ double semi_major_axis(double mtot, double P) {
    double ps = P * 86400.0;
    return pow((kG * mtot * kMsun) / (4.0 * (kPi * kPi)) * (ps * ps), 1.0 / 3.0);
}

// inc_min = arccos(Ptra)*180/pi where Ptra <= 1 else 90          marginal_likelihoods.py:120-121
__device__ __forceinline__ bool transits(double inc_deg, double Ptra) {
    double inc_min = 90.0;
    if (Ptra <= 1.0) inc_min = acos(Ptra) * 180.0 / kPi;
    return inc_deg >= inc_min;
}

// A draw whose chord is short (cos i > 0.92 Ptra, i.e. impact parameter above 0.92 (1 + k):
// at most 40 % of the central transit duration) is cheap to evaluate.  Such draws go to the back
// of the work list and are handed out last, so that the warps that finish the launch are busy
// with short items and the tail of the persistent kernel is short (longest-first scheduling
// with two classes; the result does not depend on the order).
__device__ __forceinline__ bool short_chord(double inc_deg, double Ptra) {
    return Ptra <= 1.0 && cos(inc_deg * (kPi / 180.0)) > 0.92 * Ptra;
}

__device__ __forceinline__ void push_one(unsigned ballot, bool mine, int64_t item, int64_t* slot0,
                                         int dir, unsigned long long* counter) {
    if (ballot == 0) return;
    int lane = threadIdx.x & 31;
    int leader = __ffs(ballot) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (mine) {
        long long off = (long long)(base + __popc(ballot & ((1u << lane) - 1u)));
        slot0[dir * off] = item;
    }
}

__device__ __forceinline__ void push_item(bool take, bool is_short, int64_t item, int64_t* items,
                                          int64_t cap, unsigned long long* n_items) {
    const unsigned front = __ballot_sync(0xffffffffu, take && !is_short);
    const unsigned back = __ballot_sync(0xffffffffu, take && is_short);
    push_one(front, take && !is_short, item, items, +1, n_items + 0);
    push_one(back, take && is_short, item, items + (cap - 1), -1, n_items + 3);
}

__global__ void geometry_tp_kernel(GeomTp g) {
    int64_t n_round = (g.N + 31) / 32 * 32;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_round;
         i += (int64_t)gridDim.x * blockDim.x) {
        bool take = false, brief = false;
        if (i < g.N) {
            double rp = g.rp.at(i), P = g.P.at(i), inc = g.inc.at(i), ecc = g.ecc.at(i);
            double argp = g.argp.at(i), rhost = g.rhost.at(i);
            double a = semi_major_axis(g.mtot.at(i), P);
            // e_corr, Ptra, coll                      marginal_likelihoods.py:111-115
            double e_corr = (1.0 + ecc * sin(argp * kPi / 180.0)) / (1.0 - ecc * ecc);
            double rsum = rp * kRearth + rhost * kRsun;
            double Ptra = rsum / a * e_corr;
            bool coll = rsum > a * (1.0 - ecc);
            take = transits(inc, Ptra) && !coll;
            if (g.extra_mask) take = take && (g.extra_mask[i] != 0);
            g.lnl_out[i] = neg_inf();
            if (take) {
                g.a_out[i] = a;
                brief = short_chord(inc, Ptra);
            }
            if (g.mask_out) g.mask_out[i] = take ? 1 : 0;
        }
        push_item(take, brief, i << 1, g.items, g.N, g.n_items);
    }
}

__global__ void geometry_eb_kernel(GeomEb g) {
    int64_t n_round = (g.N + 31) / 32 * 32;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_round;
         i += (int64_t)gridDim.x * blockDim.x) {
        bool take = false, brief = false;
        int twin = 0;
        if (i < g.N) {
            double reb = g.reb.at(i), q = g.q.at(i), P = g.P.at(i), inc = g.inc.at(i);
            double ecc = g.ecc.at(i), argp = g.argp.at(i), rhost = g.rhost.at(i);
            double mtot = g.mtot.at(i);
            // marginal_likelihoods.py:254-268
            double e_corr = (1.0 + ecc * sin(argp * kPi / 180.0)) / (1.0 - ecc * ecc);
            double rsum = reb * kRsun + rhost * kRsun;
            twin = (q >= 0.95) ? 1 : 0;
            double a, Ptra;
            bool coll;
            if (!twin) {
                a = semi_major_axis(mtot, P);
                Ptra = rsum / a * e_corr;
                coll = rsum > a * (1.0 - ecc);
            } else {
                a = semi_major_axis(mtot, 2.0 * P);
                Ptra = rsum / a * e_corr;
                coll = (2.0 * rhost * kRsun) > a * (1.0 - ecc);
            }
            take = transits(inc, Ptra) && !coll;
            if (g.extra_mask) take = take && (g.extra_mask[i] != 0);
            g.lnl_out[i] = neg_inf();
            g.lnl_twin_out[i] = neg_inf();
            if (take) {
                g.a_out[i] = a;
                g.p_out[i] = twin ? 2.0 * P : P;
                brief = short_chord(inc, Ptra);
            }
            if (g.mask_out) g.mask_out[i] = (take && !twin) ? 1 : 0;
            if (g.mask_twin_out) g.mask_twin_out[i] = (take && twin) ? 1 : 0;
        }
        push_item(take, brief, (i << 1) | twin, g.items, g.N, g.n_items);
        unsigned bt = __ballot_sync(0xffffffffu, take && twin);
        if ((threadIdx.x & 31) == 0 && bt) atomicAdd(g.n_items + 1, (unsigned long long)__popc(bt));
    }
}

// ---- light curve + chi^2 ---------------------------------------------------------------------
struct LnlArgs {
    LightCurve lc;
    OrbitTable tab;
    int eb;                  // 0 TP-type, 1 EB-type
    int companion_is_host;
    int raw;                 // 1: store +0.5 chi^2 (the lnL_*_p seam), +inf on the depth cut
                             // 0: store -0.5 ln(2 pi) - ln(sigma) - 0.5 chi^2 (marginal_likelihoods.py:130)
    int twin_uniform;        // twin flag when items == nullptr
    Col body;                // R_p [R_earth] (TP) or R_EB [R_sun] (EB)
    Col ebfr;                // EB flux ratio (EB only)
    Col P, inc, a, rhost, u1, u2, ecc, argp, cfr;
    const int64_t* items;    // nullptr: identity list 0..count-1
    int64_t count;             // number of work items when count_dev == nullptr
    const unsigned long long* count_dev;  // else read from device memory (written by geometry):
                                          // count_dev[0] items at the front of `items`,
                                          // count_dev[3] at the back (handed out last)
    int64_t items_cap;         // capacity of `items` (the back grows down from items_cap - 1)
    unsigned long long* next;  // work-queue cursor
    double* out;             // lnL of the (EB) branch, indexed by sample
    double* out_twin;        // lnL of the twin branch (fused EB only)
    unsigned long long* counters;  // optional [4]: stamps inside transit windows, stamps whose
                                   // sub-exposures were evaluated (window minus centre-probe
                                   // skips), interior-case points, limb/edge-case points
    // simulate mode (simulate_TP_transit_p / simulate_EB_transit_p, likelihoods.py:302-439):
    double* model_out;       // optional [count][npts] diluted model flux, caller's stamp order
    double* secdepth_out;    // optional [count] secondary-eclipse depth (EB-type)
    const int* perm;         // sorted stamp j -> caller's stamp index
    int scalar_rule;         // 1: radius-ratio rules of the scalar simulate_EB_transit
                             // (likelihoods.py:121-123, :137) instead of the vectorised ones
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

constexpr int kLnlThreads = 128;
#ifndef TRI_LNL_MIN_BLOCKS
#define TRI_LNL_MIN_BLOCKS 6       // 80 registers per thread: 24 warps per SM (A/B: 4->137, 5->130, 6->128, 7->128, 8->130 ms)
#endif
constexpr int kLnlMinBlocks = TRI_LNL_MIN_BLOCKS;

constexpr int kToffTable = 64;

__global__ void __launch_bounds__(kLnlThreads, kLnlMinBlocks) lnl_kernel(LnlArgs A) {
    extern __shared__ double smem[];
    // Stage the folded light curve once per block when it fits (else read through L1/L2).
    LightCurve lc = A.lc;
    if (A.lc.time == nullptr) return;
    {
        size_t need = (size_t)(3 * lc.npts + 1) * sizeof(double);
        unsigned dyn;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        if (need <= dyn) {
            double* st = smem;
            double* sf = smem + lc.npts;
            double* sp = smem + 2 * lc.npts;
            for (int j = threadIdx.x; j < lc.npts; j += blockDim.x) {
                st[j] = A.lc.time[j];
                sf[j] = A.lc.flux[j];
            }
            for (int j = threadIdx.x; j <= lc.npts; j += blockDim.x) sp[j] = A.lc.prefix[j];
            __syncthreads();
            lc.time = st;
            lc.flux = sf;
            lc.prefix = sp;
        }
    }
    // sub-exposure offsets of the observed light curve (the same for every draw)
    __shared__ double s_toff[kToffTable];
    const bool toff_tab = lc.nsamples < kToffTable;
    if (toff_tab) {
        const double inv = 1.0 / lc.nsamples;
        for (int is = threadIdx.x; is <= lc.nsamples; is += blockDim.x)
            s_toff[is] = is ? lc.exptime * ((is - 0.5) * inv - 0.5) : 0.0;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const double sigma = lc.sigma;
    // the Gaussian constant, once per light curve (marginal_likelihoods.py:130)
    const double lnorm = -0.5 * log(2.0 * kPi) - log(sigma);
    unsigned long long n_stamps = 0;
    unsigned n_interior = 0, n_limb = 0;   // per lane; flushed per draw
    unsigned n_skip = 0;                    // per lane: window stamps the centre probe dismissed
    unsigned long long n_int_tot = 0, n_limb_tot = 0;
    const int64_t n_front = A.count_dev ? (int64_t)A.count_dev[0] : A.count;
    const int64_t count = n_front + (A.count_dev ? (int64_t)A.count_dev[3] : 0);

#pragma unroll 1
    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(A.next, 1ull);
        w = __shfl_sync(0xffffffffu, w, 0);
        if ((int64_t)w >= count) break;
        int64_t item = !A.items ? (((int64_t)w << 1) | A.twin_uniform)
                     : ((int64_t)w < n_front ? A.items[w]
                                             : A.items[A.items_cap - 1 - ((int64_t)w - n_front)]);
        const int64_t i = item >> 1;
        const int twin = (int)(item & 1);

        // ---- per-sample constants (every lane computes the same values)
        const double P = A.P.at(i), e = A.ecc.at(i), argp = A.argp.at(i);
        const double rhost = A.rhost.at(i);
        const double a_rs = A.a.at(i) / (rhost * kRsun);          // likelihoods.py:343 / :409
        const double inc = A.inc.at(i) * (kPi / 180.0);            // :344 / :410
        const double cfr = A.cfr.at(i);
        const double F_comp = cfr / (1.0 - cfr);
        const double u1 = A.u1.at(i), u2 = A.u2.at(i);
        double k_pri, k_sec = 0.0, F_EB = 0.0;
        Dilution D;
        if (!A.eb) {
            k_pri = A.body.at(i) * kRearth / (rhost * kRsun);      // :340
            D.two_stage = false;
            D.d1 = 0.0;
            D.d2 = A.companion_is_host ? 1.0 / F_comp : F_comp / 1.0;   // :352-357
        } else {
            const double reb = A.body.at(i);
            const double fr = A.ebfr.at(i);
            F_EB = fr / (1.0 - fr);
            k_pri = reb / rhost;
            if (!A.scalar_rule) {
                if ((k_pri - 1.0) < 1e-6) k_pri *= 0.999;          // :405-406 (no abs: every k <= 1)
                k_sec = rhost / reb;                               // :417-418
                if ((k_sec - 1.0) < 1e-6) k_sec *= 0.999;
            } else {
                if (fabs(k_pri - 1.0) < 1e-6) k_pri *= 0.999;      // :121-123
                k_sec = 1.0 / k_pri;                               // :137
            }
            D.two_stage = true;
            if (A.companion_is_host) {                              // :427-432
                D.d1 = F_EB / F_comp;
                D.d2 = 1.0 / (F_comp + F_EB);
            } else {                                                // :433-438
                D.d1 = F_EB / 1.0;
                D.d2 = F_comp / (1.0 + F_EB);
            }
        }
        double* outp = (twin && A.out_twin) ? A.out_twin : A.out;

        // Two passes through ONE copy of the model code: pass 0 (EB-type only) is the
        // 25-stamp secondary eclipse on [-0.05, 0.05] d with the roles swapped and no
        // supersampling (likelihoods.py:417-423), reduced with a minimum; pass 1 is the
        // observed light curve, reduced to chi^2.
        double chi = 0.0;
        bool cut = false;
        int jlo = 0, jhi = 0;
#pragma unroll 1
        for (int pass = A.eb ? 0 : 1; pass < 2; ++pass) {
            const bool primary = (pass == 1);
            const double k = primary ? k_pri : k_sec;
            // w = (90 - argp) pi/180, + 180 deg for the secondary          :345 / :419
            const double w_rad = primary ? (90.0 - argp) * (kPi / 180.0)
                                         : (90.0 - argp + 180.0) * (kPi / 180.0);
            Orbit o;
            orbit_setup(o, A.tab, k, P, a_rs, inc, e, w_rad);
            Limb L;
            limb_setup(L, u1, u2, k);
            const int ns = primary ? lc.nsamples : 1;
            const double exptime = primary ? lc.exptime : 0.0;
            const double inv_ns = 1.0 / ns;
            if (primary) {
                // time stamps that can be in transit
                jlo = 0;
                jhi = lc.npts;
                Window win;
                if (transit_window(o, A.tab, a_rs, P, lc, win)) {
                    double half = 0.5 * lc.exptime;
                    jlo = lower_bound(lc.time, lc.npts, win.t_lo - half);
                    jhi = lower_bound(lc.time, lc.npts, win.t_hi + half);
                    if (jhi < jlo) jhi = jlo;
                }
                if (A.model_out) {   // outside the window the model is exactly 1
                    double* row = A.model_out + (size_t)w * lc.npts;
                    for (int j = lane; j < lc.npts; j += 32)
                        if (j < jlo || j >= jhi) row[A.perm[j]] = 1.0;
                }
            } else {
                jlo = 0;
                jhi = 25;
            }
            // centre probe (see max_projected_speed): only worth it with supersampling
            const bool probe = primary && ns > 1 && !o.table_clamped;
            const double skip_beyond =
                1.0 + k + max_projected_speed(o, a_rs) * (0.5 * exptime) + 1e-9;
            double red = primary ? 0.0 : INFINITY;
            // lane <-> time stamp, serial over sub-exposures.  Full rounds take 16 stamps from
            // the front of the window and 16 from its back: ingress and egress mirror each
            // other, so the two halves of the warp are in the same occultation case (limb or
            // interior) at the same time instead of one half waiting for the other.  The
            // remainder (< 32 stamps, in the middle) forms the last round, where 2 (<= 16 stamps)
            // or 4 (<= 8) lanes share one stamp's sub-exposures.
            const int n_full = (jhi - jlo) >> 5;
            const int mid_lo = jlo + 16 * n_full, rem = (jhi - jlo) - 32 * n_full;
#pragma unroll 1
            for (int r = 0; r < n_full + (rem > 0 ? 1 : 0); ++r) {
                int gsh = 0, j, sub = 0;
                bool have = true;
                bool mirror = false;
                if (r < n_full) {
                    mirror = lane >= 16;
                    j = mirror ? jhi - 16 * r - 1 - (lane - 16) : jlo + 16 * r + lane;
                } else {
                    gsh = (primary && ns >= 4) ? (rem <= 8 ? 2 : (rem <= 16 ? 1 : 0)) : 0;
                    j = mid_lo + (lane >> gsh);
                    sub = lane & ((1 << gsh) - 1);
                    have = (lane >> gsh) < rem;
                }
                double acc = 0.0;
                if (have) {
                    const double t = primary ? lc.time[j]
                                             : ((j == 24) ? 0.05 : -0.05 + j * ((0.05 - -0.05) / 24.0));
                    // this lane's sub-exposures: is_lo .. is_hi of 1 .. ns
                    const int is_lo = 1 + ((sub * ns) >> gsh), is_hi = ((sub + 1) * ns) >> gsh;
#pragma unroll 1
                    for (int is = probe ? 0 : is_lo; is <= is_hi; ++is) {
                        // sub-exposure offset exptime ((is - 1/2)/ns - 1/2), tabulated per block
                        // (the back half of a paired round runs its sub-exposures backwards in
                        // time, the mirror image of the front half)
                        const int iso = (mirror && is) ? ns + 1 - is : is;
                        const double toff = toff_tab ? (primary ? s_toff[iso] : 0.0)
                                                     : (iso ? exptime * ((iso - 0.5) * inv_ns - 0.5) : 0.0);
                        const double z = z_at(o, A.tab, t + toff);
                        if (is == 0) {   // stamp centre: is the whole exposure out of transit?
                            if (fabs(z) > skip_beyond) {
                                acc = (double)(is_hi - is_lo + 1);
                                n_skip += (sub == 0);
                                break;
                            }
                            is = is_lo - 1;
                            continue;
                        }
                        int cls = 0;   // work class of SURVEY.md 8(d): 1 interior, 2 limb-crossing
                        acc += (z > 1.0 + k) ? 1.0 : occult_quad(z, k, L, cls);
                        n_interior += (unsigned)(primary && cls == 1);
                        n_limb += (unsigned)(primary && cls == 2);
                    }
                }
                if (gsh) {   // (warp-uniform) the lanes of a stamp pool their sub-exposure sums
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    if (gsh == 2) acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                }
                if (have && sub == 0) {
                    const double m = acc / ns;
                    if (primary) {
                        const double md = dilute(D, m);
                        if (A.model_out) A.model_out[(size_t)w * lc.npts + A.perm[j]] = md;
                        const double r = lc.flux[j] - md;
                        red = fma(r, r, red);
                    } else {
                        red = fmin(red, m);
                    }
                }
            }
            if (!primary) {
                double sec = warp_min(red);
                if (A.companion_is_host) sec = (sec + F_comp / F_EB) / (1.0 + F_comp / F_EB);
                else sec = (sec + 1.0 / F_EB) / (1.0 + 1.0 / F_EB);
                const double sd = 1.0 - (sec + D.d2) / (1.0 + D.d2);
                if (A.secdepth_out && lane == 0) A.secdepth_out[w] = sd;
                cut = !twin && !(sd < 1.5 * sigma);                 // :535-538
                if (cut) break;
            } else {
                chi = warp_sum(red);
            }
        }
        if (cut) {
            if (lane == 0) outp[i] = A.raw ? INFINITY : -INFINITY;
            continue;
        }
        chi += (lc.prefix[jlo] - lc.prefix[0]) + (lc.prefix[lc.npts] - lc.prefix[jhi]);
        if (lane == 0) {
            double half_chi2 = 0.5 * (chi / (sigma * sigma));       // likelihoods.py:486
            outp[i] = A.raw ? half_chi2 : lnorm - half_chi2;
            n_stamps += (unsigned long long)(jhi - jlo);
        }
        n_int_tot += n_interior;
        n_limb_tot += n_limb;
        n_interior = n_limb = 0;
    }
    if (A.counters) {
        unsigned long long n_skip_tot = n_skip;
        for (int o = 16; o > 0; o >>= 1) {
            n_skip_tot += __shfl_xor_sync(0xffffffffu, n_skip_tot, o);
            n_int_tot += __shfl_xor_sync(0xffffffffu, n_int_tot, o);
            n_limb_tot += __shfl_xor_sync(0xffffffffu, n_limb_tot, o);
        }
        if (lane == 0) {
            atomicAdd(A.counters + 0, n_stamps);   // stamps inside the transit windows
            atomicAdd(A.counters + 1, n_stamps - n_skip_tot);
            atomicAdd(A.counters + 2, n_int_tot);
            atomicAdd(A.counters + 3, n_limb_tot);
        }
    }
}

// ---- log-mean-exp ----------------------------------------------------------------------------
struct LsePartial {
    double m;      // running max of the finite entries (-inf if none)
    double s;      // sum of exp(x - m)
    unsigned long long n_finite, n_posinf;
};

__device__ __forceinline__ void lse_push(LsePartial& a, double x) {
    if (isinf(x) && x > 0) { a.n_posinf++; return; }
    if (!isfinite(x)) return;  // -inf and NaN: zero weight
    a.n_finite++;
    if (x > a.m) {
        a.s = a.s * exp(a.m - x) + 1.0;   // exp(-inf) = 0 on the first finite entry
        a.m = x;
    } else {
        a.s += exp(x - a.m);
    }
}

__device__ __forceinline__ void lse_merge(LsePartial& a, const LsePartial& b) {
    a.n_finite += b.n_finite;
    a.n_posinf += b.n_posinf;
    if (b.m == -INFINITY) return;
    if (a.m == -INFINITY) { a.m = b.m; a.s = b.s; return; }
    if (b.m > a.m) {
        a.s = a.s * exp(a.m - b.m) + b.s;
        a.m = b.m;
    } else {
        a.s += b.s * exp(b.m - a.m);
    }
}

constexpr int kLseThreads = 256;

// one partial per block over a contiguous slice: deterministic for a given grid
__global__ void __launch_bounds__(kLseThreads)
lse_partial_kernel(const double* lnl, Col lnprior, int64_t N, LsePartial* partials) {
    __shared__ LsePartial sh[kLseThreads];
    int64_t per_block = (N + gridDim.x - 1) / gridDim.x;
    int64_t lo = blockIdx.x * per_block;
    int64_t hi = lo + per_block < N ? lo + per_block : N;
    LsePartial a{-INFINITY, 0.0, 0, 0};
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        double x = lnl[i];
        if (lnprior.p) x += lnprior.at(i);
        lse_push(a, x);
    }
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = kLseThreads / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) lse_merge(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = sh[0];
}

__global__ void lse_final_kernel(const LsePartial* partials, int n, LsePartial* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    LsePartial a{-INFINITY, 0.0, 0, 0};
    for (int b = 0; b < n; ++b) lse_merge(a, partials[b]);
    *out = a;
}

// ---- best draws: top-K of lnL on the device (reference marginal_likelihoods.py:152-153) --------
// Radix select on the order-preserving integer image of the doubles: eight 8-bit passes find the
// exact K-th largest finite value T; everything above T is appended by the whole grid, ties at
// T are taken in index order by one block, so the selection equals a stable sort by
// (-lnL, index).  NaN and -inf never qualify.
struct TopkState {
    unsigned long long prefix, mask;   // key bits fixed so far
    unsigned long long k_remaining;    // rank still to resolve inside the current bucket
    unsigned long long n_finite;       // finite entries seen in pass 0
    unsigned long long n_out;          // entries written to the output
    unsigned long long n_ties;         // tie slots to fill with keys == prefix
    unsigned long long n_equal;        // entries whose key == prefix
    unsigned int hist[256];
};

__device__ __forceinline__ unsigned long long topk_key(double x) {
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);   // larger double -> larger key
}

__device__ __forceinline__ bool topk_valid(double x) { return isfinite(x); }

constexpr int kTopkThreads = 256;

__global__ void topk_init_kernel(TopkState* st, unsigned long long k) {
    if (threadIdx.x < 256) st->hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        st->prefix = 0; st->mask = 0; st->k_remaining = k; st->n_finite = 0; st->n_out = 0;
        st->n_ties = 0; st->n_equal = 0;
    }
}

__global__ void __launch_bounds__(kTopkThreads)
topk_hist_kernel(const double* lnl, int64_t N, TopkState* st, int pass) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = st->prefix, mask = st->mask;
    const int shift = 56 - 8 * pass;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * blockDim.x) {
        double x = lnl[i];
        if (!topk_valid(x)) continue;
        unsigned long long key = topk_key(x);
        if ((key & mask) == prefix) atomicAdd(&h[(key >> shift) & 0xff], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void topk_scan_kernel(TopkState* st, int pass) {
    if (threadIdx.x != 0) return;
    const int shift = 56 - 8 * pass;
    unsigned long long total = 0;
    for (int b = 0; b < 256; ++b) total += st->hist[b];
    if (pass == 0) {
        st->n_finite = total;
        if (st->k_remaining > total) st->k_remaining = total;   // fewer finite entries than K
    }
    unsigned long long need = st->k_remaining, above = 0, in_bucket = 0;
    int bucket = 0;
    for (int b = 255; b >= 0; --b) {
        unsigned long long c = st->hist[b];
        if (above + c >= need && c > 0) { bucket = b; in_bucket = c; break; }
        above += c;
    }
    if (need == 0) bucket = 255;
    st->prefix |= (unsigned long long)bucket << shift;
    st->mask |= 0xffull << shift;
    st->k_remaining = need - above;     // rank inside the chosen bucket
    if (pass == 7) {
        st->n_ties = st->k_remaining;   // how many entries equal to T are wanted ...
        st->n_equal = in_bucket;        // ... out of this many
    }
    for (int b = 0; b < 256; ++b) st->hist[b] = 0;
}

// entries strictly above the threshold (and the ones equal to it when all of them are wanted,
// the usual case of a unique K-th value): order does not matter, they are sorted afterwards
__global__ void __launch_bounds__(kTopkThreads)
topk_collect_above_kernel(const double* lnl, int64_t N, TopkState* st, int64_t* out_idx,
                          double* out_val, int64_t cap) {
    if (st->n_finite == 0) return;
    const unsigned long long T = st->prefix;
    const bool all_ties = st->n_equal <= st->n_ties;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * blockDim.x) {
        double x = lnl[i];
        if (!topk_valid(x)) continue;
        unsigned long long key = topk_key(x);
        if (key < T || (key == T && !all_ties)) continue;
        unsigned long long pos = atomicAdd(&st->n_out, 1ull);
        if ((int64_t)pos < cap) { out_idx[pos] = i; out_val[pos] = x; }
    }
}

// entries equal to the threshold, lowest indices first (one block walks the array in order)
__global__ void __launch_bounds__(1024)
topk_collect_ties_kernel(const double* lnl, int64_t N, TopkState* st, int64_t* out_idx,
                         double* out_val, int64_t cap) {
    __shared__ unsigned int warp_cnt[32];
    __shared__ unsigned long long base_sh;
    __shared__ long long want_sh;
    if (st->n_finite == 0 || st->n_equal <= st->n_ties) return;   // nothing left to order
    const unsigned long long T = st->prefix;
    if (threadIdx.x == 0) { base_sh = st->n_out; want_sh = (long long)st->n_ties; }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t start = 0; start < N; start += blockDim.x) {
        long long want = want_sh;
        if (want <= 0) break;
        int64_t i = start + threadIdx.x;
        bool hit = false;
        double x = 0.0;
        if (i < N) { x = lnl[i]; hit = topk_valid(x) && topk_key(x) == T; }
        unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            unsigned c = warp_cnt[w];
            if (w < wid) before += c;
            total += c;
        }
        unsigned rank = before + __popc(bal & ((1u << lane) - 1u));
        if (hit && (long long)rank < want) {
            unsigned long long pos = base_sh + rank;
            if ((int64_t)pos < cap) { out_idx[pos] = i; out_val[pos] = x; }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned take = total < (unsigned long long)want ? total : (unsigned)want;
            base_sh += take;
            want_sh = want - take;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) st->n_out = base_sh;
}

// ---- FITPACK B-spline evaluation (device sampler: stellar relations, funcs.py:54-140) ----------
// y[i] = s(x[i]) for the spline (t[n], c[n], degree k <= 5) with extrapolation from the end
// intervals (splev ext=0): interval search + de Boor's recurrence (fpbspl), one thread per value.
constexpr int kSplevMaxKnots = 512;

template <int K>
__global__ void __launch_bounds__(256) splev_kernel(const double* __restrict__ t,
                                                     const double* __restrict__ c, int n,
                                                     const double* x, double* y, int64_t N) {
    constexpr int k = K;
    __shared__ double st[kSplevMaxKnots], sc[kSplevMaxKnots];
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        st[j] = t[j];
        sc[j] = c[j];
    }
    __syncthreads();
    const int lmin = k, lmax = n - k - 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double xv = x[i];
        // l = (number of knots <= x) - 1, clamped to the interior intervals
        int lo = 0, hi = n;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (st[mid] <= xv) lo = mid + 1; else hi = mid;
        }
        int l = min(max(lo - 1, lmin), lmax);
        double h[K + 1], hh[K > 0 ? K : 1];
        h[0] = 1.0;
#pragma unroll
        for (int j = 1; j <= k; ++j) {
#pragma unroll
            for (int q = 0; q < j; ++q) hh[q] = h[q];
            h[0] = 0.0;
#pragma unroll
            for (int q = 0; q < j; ++q) {
                const double ti = st[l + q + 1], tj = st[l + q + 1 - j];
                const double f = hh[q] / (ti - tj);
                h[q] = h[q] + f * (ti - xv);
                h[q + 1] = f * (xv - tj);
            }
        }
        double sp = 0.0;
#pragma unroll
        for (int j = 0; j <= k; ++j) sp += sc[l - k + j] * h[j];
        y[i] = sp;
    }
}

// ---- FP64 issue-rate probe (roofline denominator measured on the box) -----------------------
__global__ void dfma_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double x = 1.0000001, y = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
        a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace tri
