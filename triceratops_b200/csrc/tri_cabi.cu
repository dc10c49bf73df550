// C ABI of the engine (include/triceratops_b200.h): device context, light-curve upload,
// staging of host columns, kernel launches, result read-back.  No CPU compute path: every entry
// point fails with TRI_ENODEVICE / TRI_ECUDA when there is no usable GPU.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/triceratops_b200.h"
#include "tri_kernels.cuh"

namespace {

using namespace tri;

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(TRI_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// grow-only device arena, bump-allocated per call
struct Arena {
    char* base = nullptr;
    size_t cap = 0, used = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return TRI_OK;
        if (base) CU(cudaFree(base));
        base = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4;
        CU(cudaMalloc(&base, want));
        cap = want;
        return TRI_OK;
    }
    void reset() { used = 0; }
    template <typename T>
    T* take(size_t n) {
        size_t off = (used + 255) & ~size_t(255);
        used = off + n * sizeof(T);
        return reinterpret_cast<T*>(base + off);
    }
};

struct Ctx {
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    int lnl_blocks_per_sm = 0;
    size_t lnl_smem = 0;   // dynamic shared memory used to stage the light curve (0 = none)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool timing_valid = false;
    int launches = 0;
    double* d_tae = nullptr;
    OrbitTable tab{};
    // light curve
    bool have_lc = false;
    double *d_time = nullptr, *d_flux = nullptr, *d_prefix = nullptr;
    int* d_perm = nullptr;   // sorted stamp -> caller's stamp
    size_t lc_cap = 0;
    LightCurve lc{};
    Arena scratch;   // kernel scratch (a, p, lnl, items, partials)
    Arena staging;   // device copies of host columns
    unsigned long long* d_counters = nullptr;  // [8]
    LsePartial* d_lse_out = nullptr;           // [2]
    LsePartial* h_lse_out = nullptr;           // pinned [2]
    unsigned long long* h_counters = nullptr;  // pinned [8]
    TopkState* d_topk = nullptr;               // [2]
    TopkState* h_topk = nullptr;               // pinned [2]
    const double* last_lnl[2] = {nullptr, nullptr};   // device lnL arrays of the last eval
    int64_t last_N = 0;
};

Ctx g;

// ---- orbit table (same construction as oracle/quadmodel.py make_orbit_table) ---------------
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

void linspace(double start, double stop, int n, std::vector<double>& out) {
    out.resize(n);
    double step = (stop - start) / (double)(n - 1);
    for (int i = 0; i < n; i++) out[i] = start + (double)i * step;
    out[n - 1] = stop;
}

int need_ready(bool need_lc) {
    if (!g.ready) return fail(TRI_ESTATE, "tri_init has not been called");
    if (need_lc && !g.have_lc) return fail(TRI_ESTATE, "tri_set_lightcurve has not been called");
    return TRI_OK;
}

Col to_col(const tri_col& c) { return Col{c.ptr, c.stride}; }

// copy one host column to the staging arena
int stage(const tri_col& h, int64_t N, Col& d, cudaStream_t s) {
    if (h.ptr == nullptr) {
        d = Col{nullptr, 0};
        return TRI_OK;
    }
    if (h.stride != 0 && h.stride != 1) return fail(TRI_EINVAL, "column stride must be 0 or 1");
    size_t n = h.stride ? (size_t)N : 1;
    double* p = g.staging.take<double>(n);
    CU(cudaMemcpyAsync(p, h.ptr, n * sizeof(double), cudaMemcpyHostToDevice, s));
    d = Col{p, h.stride};
    return TRI_OK;
}

int lnl_grid() { return g.sm_count * std::max(1, g.lnl_blocks_per_sm); }

size_t lnl_smem_bytes() { return g.lnl_smem; }

struct Scratch {
    double *a = nullptr, *p = nullptr, *lnl = nullptr, *lnl_twin = nullptr;
    int64_t* items = nullptr;
    uint8_t *mask = nullptr, *mask_twin = nullptr;
    LsePartial* partials = nullptr;
    int n_partials = 0;
};

int lse_blocks(int64_t N) {
    int64_t b = (N + 4095) / 4096;
    return (int)std::max<int64_t>(1, std::min<int64_t>(b, 4 * (int64_t)g.sm_count));
}

size_t scratch_bytes(int64_t N, bool eb) {
    size_t n = (size_t)N;
    size_t b = 0;
    b += (n * 8 + 256) * (eb ? 4 : 2);   // a, lnl (+ p, lnl_twin)
    b += n * 8 + 256;                    // items
    b += (n + 256) * 2;                  // masks
    b += (size_t)lse_blocks(N) * sizeof(LsePartial) * 2 + 512;
    return b + 4096;
}

void finish_result(const LsePartial& r, int64_t N, tri_result* out) {
    out->m = r.m;
    out->s = r.s;
    out->n_finite = (int64_t)r.n_finite;
    out->n_posinf = (int64_t)r.n_posinf;
    if (r.n_posinf > 0) out->lnZ = INFINITY;
    else if (r.n_finite == 0) out->lnZ = -INFINITY;
    else out->lnZ = r.m + std::log(r.s) - std::log((double)N);
}

int launch_lse(const double* lnl, Col lnprior, int64_t N, LsePartial* partials, LsePartial* out,
               cudaStream_t s) {
    int nb = lse_blocks(N);
    lse_partial_kernel<<<nb, kLseThreads, 0, s>>>(lnl, lnprior, N, partials);
    lse_final_kernel<<<1, 32, 0, s>>>(partials, nb, out);
    g.launches += 2;
    CU(cudaGetLastError());
    return TRI_OK;
}

int launch_lnl(LnlArgs& A, cudaStream_t s) {
    size_t smem = lnl_smem_bytes();
    lnl_kernel<<<lnl_grid(), kLnlThreads, smem, s>>>(A);
    g.launches += 1;
    CU(cudaGetLastError());
    return TRI_OK;
}

// top-K of lnl on the device into (d_idx, d_val); the state record of `slot` receives n_out
int launch_topk(const double* lnl, int64_t N, int64_t cap, int slot, int64_t* d_idx,
                double* d_val, cudaStream_t s) {
    TopkState* st = g.d_topk + slot;
    int nb = (int)std::max<int64_t>(1, std::min<int64_t>((N + 4095) / 4096,
                                                         4 * (int64_t)g.sm_count));
    topk_init_kernel<<<1, 256, 0, s>>>(st, (unsigned long long)cap);
    for (int pass = 0; pass < 8; ++pass) {
        topk_hist_kernel<<<nb, kTopkThreads, 0, s>>>(lnl, N, st, pass);
        topk_scan_kernel<<<1, 32, 0, s>>>(st, pass);
    }
    topk_collect_above_kernel<<<nb, kTopkThreads, 0, s>>>(lnl, N, st, d_idx, d_val, cap);
    topk_collect_ties_kernel<<<1, 1024, 0, s>>>(lnl, N, st, d_idx, d_val, cap);
    g.launches += 19;
    CU(cudaGetLastError());
    return TRI_OK;
}

// core of tri_eval_tp*: every pointer in `a` is a device pointer
int eval_tp_device(const tri_tp_args& a, tri_result* out, cudaStream_t s) {
    const int64_t N = a.N;
    g.scratch.reset();
    int rc = g.scratch.reserve(scratch_bytes(N, false));
    if (rc) return rc;
    Scratch S;
    S.a = g.scratch.take<double>(N);
    S.lnl = out->lnL_out ? out->lnL_out : g.scratch.take<double>(N);
    S.items = g.scratch.take<int64_t>(N);
    S.partials = g.scratch.take<LsePartial>(lse_blocks(N));
    g.launches = 0;
    CU(cudaMemsetAsync(g.d_counters, 0, 8 * sizeof(unsigned long long), s));
    CU(cudaEventRecord(g.ev[0], s));
    GeomTp G{};
    G.N = N;
    G.rp = to_col(a.rp); G.P = to_col(a.P_orb); G.inc = to_col(a.inc); G.ecc = to_col(a.ecc);
    G.argp = to_col(a.argp); G.mtot = to_col(a.mtot); G.rhost = to_col(a.rhost);
    G.extra_mask = a.extra_mask;
    G.a_out = S.a; G.lnl_out = S.lnl; G.mask_out = out->mask_out;
    G.items = S.items; G.n_items = g.d_counters + 0;
    int gb = (int)std::min<int64_t>((N + 255) / 256, (int64_t)g.sm_count * 8);
    geometry_tp_kernel<<<std::max(gb, 1), 256, 0, s>>>(G);
    g.launches += 1;
    CU(cudaGetLastError());
    CU(cudaEventRecord(g.ev[1], s));

    // the work-list length stays on the device: no host round trip between the two kernels
    LnlArgs A{};
    A.lc = g.lc; A.tab = g.tab; A.eb = 0; A.companion_is_host = a.companion_is_host; A.raw = 0;
    A.twin_uniform = 0;
    A.body = to_col(a.rp); A.ebfr = Col{nullptr, 0};
    A.P = to_col(a.P_orb); A.inc = to_col(a.inc); A.a = Col{S.a, 1}; A.rhost = to_col(a.rhost);
    A.u1 = to_col(a.u1); A.u2 = to_col(a.u2); A.ecc = to_col(a.ecc); A.argp = to_col(a.argp);
    A.cfr = to_col(a.cfr);
    A.items = S.items; A.count = 0; A.count_dev = g.d_counters + 0; A.next = g.d_counters + 2;
    A.out = S.lnl; A.out_twin = nullptr; A.counters = g.d_counters + 4;
    rc = launch_lnl(A, s);
    if (rc) return rc;
    CU(cudaEventRecord(g.ev[2], s));
    rc = launch_lse(S.lnl, to_col(a.lnprior), N, S.partials, g.d_lse_out, s);
    if (rc) return rc;
    const bool want_top = out->top_cap > 0 && out->top_idx && out->top_lnL;
    if (want_top) {
        rc = launch_topk(S.lnl, N, out->top_cap, 0, out->top_idx, out->top_lnL, s);
        if (rc) return rc;
        CU(cudaMemcpyAsync(g.h_topk, g.d_topk, sizeof(TopkState), cudaMemcpyDeviceToHost, s));
    }
    g.last_lnl[0] = S.lnl; g.last_lnl[1] = nullptr; g.last_N = N;
    CU(cudaEventRecord(g.ev[3], s));
    CU(cudaMemcpyAsync(g.h_lse_out, g.d_lse_out, sizeof(LsePartial), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(g.h_counters, g.d_counters, 8 * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    g.timing_valid = true;
    finish_result(g.h_lse_out[0], N, out);
    out->n_pass = (int64_t)g.h_counters[0];
    out->n_stamps = (int64_t)g.h_counters[5];
    out->n_interior = (int64_t)g.h_counters[6];
    out->n_limb = (int64_t)g.h_counters[7];
    out->n_top = want_top ? (int64_t)std::min<unsigned long long>(g.h_topk[0].n_out,
                                                                 (unsigned long long)out->top_cap)
                          : 0;
    out->n_evaluated = want_top ? (int64_t)g.h_topk[0].n_finite : -1;
    return TRI_OK;
}

int eval_eb_device(const tri_eb_args& a, tri_result out[2], cudaStream_t s) {
    const int64_t N = a.N;
    g.scratch.reset();
    int rc = g.scratch.reserve(scratch_bytes(N, true));
    if (rc) return rc;
    Scratch S;
    S.a = g.scratch.take<double>(N);
    S.p = g.scratch.take<double>(N);
    S.lnl = out[0].lnL_out ? out[0].lnL_out : g.scratch.take<double>(N);
    S.lnl_twin = out[1].lnL_out ? out[1].lnL_out : g.scratch.take<double>(N);
    S.items = g.scratch.take<int64_t>(N);
    S.n_partials = lse_blocks(N);
    S.partials = g.scratch.take<LsePartial>(2 * S.n_partials);
    g.launches = 0;
    CU(cudaMemsetAsync(g.d_counters, 0, 8 * sizeof(unsigned long long), s));
    CU(cudaEventRecord(g.ev[0], s));
    GeomEb G{};
    G.N = N;
    G.reb = to_col(a.reb); G.q = to_col(a.q); G.P = to_col(a.P_orb); G.inc = to_col(a.inc);
    G.ecc = to_col(a.ecc); G.argp = to_col(a.argp); G.mtot = to_col(a.mtot);
    G.rhost = to_col(a.rhost);
    G.extra_mask = a.extra_mask;
    G.a_out = S.a; G.p_out = S.p; G.lnl_out = S.lnl; G.lnl_twin_out = S.lnl_twin;
    G.mask_out = out[0].mask_out; G.mask_twin_out = out[1].mask_out;
    G.items = S.items; G.n_items = g.d_counters + 0;
    int gb = (int)std::min<int64_t>((N + 255) / 256, (int64_t)g.sm_count * 8);
    geometry_eb_kernel<<<std::max(gb, 1), 256, 0, s>>>(G);
    g.launches += 1;
    CU(cudaGetLastError());
    CU(cudaEventRecord(g.ev[1], s));

    LnlArgs A{};
    A.lc = g.lc; A.tab = g.tab; A.eb = 1; A.companion_is_host = a.companion_is_host; A.raw = 0;
    A.twin_uniform = 0;
    A.body = to_col(a.reb); A.ebfr = to_col(a.ebfr);
    A.P = Col{S.p, 1}; A.inc = to_col(a.inc); A.a = Col{S.a, 1}; A.rhost = to_col(a.rhost);
    A.u1 = to_col(a.u1); A.u2 = to_col(a.u2); A.ecc = to_col(a.ecc); A.argp = to_col(a.argp);
    A.cfr = to_col(a.cfr);
    A.items = S.items; A.count = 0; A.count_dev = g.d_counters + 0; A.next = g.d_counters + 2;
    A.out = S.lnl; A.out_twin = S.lnl_twin; A.counters = g.d_counters + 4;
    rc = launch_lnl(A, s);
    if (rc) return rc;
    CU(cudaEventRecord(g.ev[2], s));
    rc = launch_lse(S.lnl, to_col(a.lnprior), N, S.partials, g.d_lse_out, s);
    if (rc) return rc;
    rc = launch_lse(S.lnl_twin, to_col(a.lnprior), N, S.partials + S.n_partials,
                    g.d_lse_out + 1, s);
    if (rc) return rc;
    bool want_top[2];
    for (int b = 0; b < 2; ++b) {
        want_top[b] = out[b].top_cap > 0 && out[b].top_idx && out[b].top_lnL;
        if (want_top[b]) {
            rc = launch_topk(b ? S.lnl_twin : S.lnl, N, out[b].top_cap, b, out[b].top_idx,
                             out[b].top_lnL, s);
            if (rc) return rc;
        }
    }
    if (want_top[0] || want_top[1])
        CU(cudaMemcpyAsync(g.h_topk, g.d_topk, 2 * sizeof(TopkState), cudaMemcpyDeviceToHost, s));
    g.last_lnl[0] = S.lnl; g.last_lnl[1] = S.lnl_twin; g.last_N = N;
    CU(cudaEventRecord(g.ev[3], s));
    CU(cudaMemcpyAsync(g.h_lse_out, g.d_lse_out, 2 * sizeof(LsePartial), cudaMemcpyDeviceToHost,
                       s));
    CU(cudaMemcpyAsync(g.h_counters, g.d_counters, 8 * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    g.timing_valid = true;
    finish_result(g.h_lse_out[0], N, &out[0]);
    finish_result(g.h_lse_out[1], N, &out[1]);
    out[0].n_pass = (int64_t)g.h_counters[0] - (int64_t)g.h_counters[1];
    out[1].n_pass = (int64_t)g.h_counters[1];
    for (int b = 0; b < 2; ++b) {   // the two branches share one launch: totals are joint
        out[b].n_stamps = (int64_t)g.h_counters[5];
        out[b].n_interior = (int64_t)g.h_counters[6];
        out[b].n_limb = (int64_t)g.h_counters[7];
        out[b].n_top = want_top[b]
            ? (int64_t)std::min<unsigned long long>(g.h_topk[b].n_out,
                                                    (unsigned long long)out[b].top_cap)
            : 0;
        out[b].n_evaluated = want_top[b] ? (int64_t)g.h_topk[b].n_finite : -1;
    }
    return TRI_OK;
}

// best first, ties by ascending index: the order of a stable sort of -lnL
void sort_candidates(int64_t n, int64_t* idx, double* val) {
    std::vector<std::pair<double, int64_t>> v((size_t)n);
    for (int64_t i = 0; i < n; ++i) v[(size_t)i] = {val[i], idx[i]};
    std::sort(v.begin(), v.end(), [](const std::pair<double, int64_t>& x,
                                     const std::pair<double, int64_t>& y) {
        return x.first > y.first || (x.first == y.first && x.second < y.second);
    });
    for (int64_t i = 0; i < n; ++i) { val[i] = v[(size_t)i].first; idx[i] = v[(size_t)i].second; }
}

int check_cols(int64_t N, std::initializer_list<const tri_col*> req) {
    if (N < 0) return fail(TRI_EINVAL, "N must be >= 0");
    for (const tri_col* c : req) {
        if (c->ptr == nullptr) return fail(TRI_EINVAL, "a required column is NULL");
        if (c->stride != 0 && c->stride != 1)
            return fail(TRI_EINVAL, "column stride must be 0 or 1");
    }
    return TRI_OK;
}

}  // namespace

extern "C" {

const char* tri_last_error(void) { return g_err.c_str(); }

int tri_init(int device) {
    if (g.ready && g.device == device) return TRI_OK;
    if (g.ready) tri_shutdown();
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(TRI_ENODEVICE,
                    std::string("no CUDA device available (there is no CPU fallback): ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= n) return fail(TRI_EINVAL, "device index out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    g.device = device;
    g.sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    for (auto& ev : g.ev) CU(cudaEventCreate(&ev));
    // orbit table
    std::vector<double> es, ms, tae((size_t)kTableNe * kTableNm);
    linspace(0.0, kTableMaxE, kTableNe, es);
    linspace(0.0, kPi, kTableNm, ms);
    for (int i = 0; i < kTableNe; i++)
        for (int j = 0; j < kTableNm; j++)
            tae[(size_t)i * kTableNm + j] = ta_newton(ms[j], es[i]) - ms[j];
    CU(cudaMalloc(&g.d_tae, tae.size() * sizeof(double)));
    CU(cudaMemcpy(g.d_tae, tae.data(), tae.size() * sizeof(double), cudaMemcpyHostToDevice));
    g.tab.tae = g.d_tae;
    g.tab.de = es[1] - es[0];
    g.tab.dm = ms[1] - ms[0];
    g.tab.inv_dm = 1.0 / g.tab.dm;
    CU(cudaMalloc(&g.d_counters, 8 * sizeof(unsigned long long)));
    CU(cudaMalloc(&g.d_lse_out, 2 * sizeof(LsePartial)));
    CU(cudaMallocHost(&g.h_lse_out, 2 * sizeof(LsePartial)));
    CU(cudaMallocHost(&g.h_counters, 8 * sizeof(unsigned long long)));
    CU(cudaMalloc(&g.d_topk, 2 * sizeof(TopkState)));
    CU(cudaMallocHost(&g.h_topk, 2 * sizeof(TopkState)));
    // occupancy of the persistent light-curve kernel
    int bps = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, lnl_kernel, kLnlThreads, 0));
    g.lnl_blocks_per_sm = std::max(1, bps);
    g.lnl_smem = 0;
    g.ready = true;
    g.have_lc = false;
    return TRI_OK;
}

int tri_shutdown(void) {
    if (!g.ready) return TRI_OK;
    cudaSetDevice(g.device);
    cudaStreamSynchronize(g.stream);
    cudaFree(g.d_tae);
    cudaFree(g.d_time);
    cudaFree(g.d_flux);
    cudaFree(g.d_prefix);
    cudaFree(g.d_perm);
    cudaFree(g.d_counters);
    cudaFree(g.d_lse_out);
    cudaFreeHost(g.h_lse_out);
    cudaFreeHost(g.h_counters);
    cudaFree(g.d_topk);
    cudaFreeHost(g.h_topk);
    if (g.scratch.base) cudaFree(g.scratch.base);
    if (g.staging.base) cudaFree(g.staging.base);
    for (auto& ev : g.ev) cudaEventDestroy(ev);
    cudaStreamDestroy(g.stream);
    g = Ctx{};
    return TRI_OK;
}

int tri_sm_count(int32_t* n) {
    int rc = need_ready(false);
    if (rc) return rc;
    *n = g.sm_count;
    return TRI_OK;
}

int tri_set_lightcurve(const double* time, const double* flux, int64_t npts, double sigma,
                       double exptime, int32_t nsamples) {
    int rc = need_ready(false);
    if (rc) return rc;
    if (!time || !flux || npts <= 0) return fail(TRI_EINVAL, "empty light curve");
    if (npts > (int64_t)1 << 28) return fail(TRI_EINVAL, "light curve too long");
    if (!(sigma > 0.0) || nsamples < 1 || !(exptime >= 0.0))
        return fail(TRI_EINVAL, "sigma must be > 0, nsamples >= 1, exptime >= 0");
    for (int64_t j = 0; j < npts; j++)
        if (std::isnan(time[j]) || std::isnan(flux[j]))
            return fail(TRI_EINVAL, "NaN in light curve (calc_probs drops them, triceratops.py:709)");
    // chi^2 is a plain sum over stamps, so the stamps may be visited in time order
    std::vector<int64_t> order(npts);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](int64_t x, int64_t y) { return time[x] < time[y]; });
    std::vector<double> t(npts), f(npts), pre(npts + 1);
    long double run = 0.0L;
    pre[0] = 0.0;
    for (int64_t j = 0; j < npts; j++) {
        t[j] = time[order[j]];
        f[j] = flux[order[j]];
        long double d = (long double)f[j] - 1.0L;
        run += (long double)((double)d * (double)d);
        pre[j + 1] = (double)run;
    }
    CU(cudaSetDevice(g.device));
    CU(cudaStreamSynchronize(g.stream));
    if ((size_t)npts > g.lc_cap) {
        cudaFree(g.d_time); cudaFree(g.d_flux); cudaFree(g.d_prefix); cudaFree(g.d_perm);
        g.d_time = g.d_flux = g.d_prefix = nullptr;
        g.d_perm = nullptr;
        g.lc_cap = 0;
        CU(cudaMalloc(&g.d_perm, npts * sizeof(int)));
        CU(cudaMalloc(&g.d_time, npts * sizeof(double)));
        CU(cudaMalloc(&g.d_flux, npts * sizeof(double)));
        CU(cudaMalloc(&g.d_prefix, (npts + 1) * sizeof(double)));
        g.lc_cap = (size_t)npts;
    }
    CU(cudaMemcpy(g.d_time, t.data(), npts * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(g.d_flux, f.data(), npts * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(g.d_prefix, pre.data(), (npts + 1) * sizeof(double), cudaMemcpyHostToDevice));
    {
        std::vector<int> perm(npts);
        for (int64_t j = 0; j < npts; j++) perm[j] = (int)order[j];
        CU(cudaMemcpy(g.d_perm, perm.data(), npts * sizeof(int), cudaMemcpyHostToDevice));
    }
    g.lc.time = g.d_time; g.lc.flux = g.d_flux; g.lc.prefix = g.d_prefix;
    g.lc.npts = (int)npts; g.lc.nsamples = nsamples; g.lc.sigma = sigma; g.lc.exptime = exptime;
    g.lc.tmin = t.front(); g.lc.tmax = t.back();
    // stage the light curve in shared memory only when that costs no occupancy
    {
        int bps0 = 0, bps1 = 0;
        size_t need = (size_t)(3 * (size_t)npts + 1) * sizeof(double);
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps0, lnl_kernel, kLnlThreads, 0));
        g.lnl_smem = 0;
        if (need <= 48 * 1024) {
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps1, lnl_kernel, kLnlThreads, need));
            if (bps1 >= bps0) g.lnl_smem = need;
        }
        g.lnl_blocks_per_sm = std::max(1, bps0);
    }
    g.have_lc = true;
    return TRI_OK;
}

int tri_eval_tp_dev(const tri_tp_args* a, tri_result* out, void* stream) {
    int rc = need_ready(true);
    if (rc) return rc;
    if (!a || !out) return fail(TRI_EINVAL, "NULL argument");
    rc = check_cols(a->N, {&a->rp, &a->P_orb, &a->inc, &a->ecc, &a->argp, &a->mtot, &a->rhost,
                           &a->u1, &a->u2, &a->cfr});
    if (rc) return rc;
    CU(cudaSetDevice(g.device));
    if (a->N == 0) {
        LsePartial z{-INFINITY, 0.0, 0, 0};
        finish_result(z, 0, out);
        out->n_pass = out->n_stamps = out->n_interior = out->n_limb = out->n_top = 0;
        return TRI_OK;
    }
    return eval_tp_device(*a, out, stream ? (cudaStream_t)stream : g.stream);
}

int tri_eval_eb_dev(const tri_eb_args* a, tri_result out[2], void* stream) {
    int rc = need_ready(true);
    if (rc) return rc;
    if (!a || !out) return fail(TRI_EINVAL, "NULL argument");
    rc = check_cols(a->N, {&a->reb, &a->ebfr, &a->q, &a->P_orb, &a->inc, &a->ecc, &a->argp,
                           &a->mtot, &a->rhost, &a->u1, &a->u2, &a->cfr});
    if (rc) return rc;
    CU(cudaSetDevice(g.device));
    if (a->N == 0) {
        LsePartial z{-INFINITY, 0.0, 0, 0};
        for (int b = 0; b < 2; ++b) {
            finish_result(z, 0, &out[b]);
            out[b].n_pass = out[b].n_stamps = out[b].n_interior = out[b].n_limb = out[b].n_top = 0;
        }
        return TRI_OK;
    }
    return eval_eb_device(*a, out, stream ? (cudaStream_t)stream : g.stream);
}

int tri_eval_tp(const tri_tp_args* a, tri_result* out) {
    int rc = need_ready(true);
    if (rc) return rc;
    if (!a || !out) return fail(TRI_EINVAL, "NULL argument");
    rc = check_cols(a->N, {&a->rp, &a->P_orb, &a->inc, &a->ecc, &a->argp, &a->mtot, &a->rhost,
                           &a->u1, &a->u2, &a->cfr});
    if (rc) return rc;
    if (a->N == 0) return tri_eval_tp_dev(a, out, nullptr);
    CU(cudaSetDevice(g.device));
    const int64_t N = a->N;
    cudaStream_t s = g.stream;
    g.staging.reset();
    rc = g.staging.reserve((size_t)N * 8 * 13 + (size_t)N * 2 + 8192
                           + (size_t)std::max<int64_t>(out->top_cap, 0) * 16);
    if (rc) return rc;
    tri_tp_args d = *a;
    Col c;
#define STAGE(field)                                   \
    rc = stage(a->field, N, c, s);                     \
    if (rc) return rc;                                 \
    d.field = tri_col{c.p, c.stride};
    STAGE(rp) STAGE(P_orb) STAGE(inc) STAGE(ecc) STAGE(argp) STAGE(mtot) STAGE(rhost)
    STAGE(u1) STAGE(u2) STAGE(cfr) STAGE(lnprior)
    if (a->extra_mask) {
        uint8_t* m = g.staging.take<uint8_t>(N);
        CU(cudaMemcpyAsync(m, a->extra_mask, (size_t)N, cudaMemcpyHostToDevice, s));
        d.extra_mask = m;
    }
    tri_result r = *out;
    double* h_lnl = out->lnL_out;
    uint8_t* h_mask = out->mask_out;
    int64_t* h_tidx = out->top_idx;
    double* h_tval = out->top_lnL;
    const bool top = out->top_cap > 0 && h_tidx && h_tval;
    r.lnL_out = h_lnl ? g.staging.take<double>(N) : nullptr;
    r.mask_out = h_mask ? g.staging.take<uint8_t>(N) : nullptr;
    r.top_cap = top ? out->top_cap : 0;
    r.top_idx = top ? g.staging.take<int64_t>(out->top_cap) : nullptr;
    r.top_lnL = top ? g.staging.take<double>(out->top_cap) : nullptr;
    rc = eval_tp_device(d, &r, s);
    if (rc) return rc;
    if (h_lnl) CU(cudaMemcpyAsync(h_lnl, r.lnL_out, (size_t)N * 8, cudaMemcpyDeviceToHost, s));
    if (h_mask) CU(cudaMemcpyAsync(h_mask, r.mask_out, (size_t)N, cudaMemcpyDeviceToHost, s));
    if (top && r.n_top > 0) {
        CU(cudaMemcpyAsync(h_tidx, r.top_idx, (size_t)r.n_top * 8, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(h_tval, r.top_lnL, (size_t)r.n_top * 8, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    if (top) sort_candidates(r.n_top, h_tidx, h_tval);
    r.lnL_out = h_lnl;
    r.mask_out = h_mask;
    r.top_cap = out->top_cap;
    r.top_idx = h_tidx;
    r.top_lnL = h_tval;
    *out = r;
    return TRI_OK;
}

int tri_eval_eb(const tri_eb_args* a, tri_result out[2]) {
    int rc = need_ready(true);
    if (rc) return rc;
    if (!a || !out) return fail(TRI_EINVAL, "NULL argument");
    rc = check_cols(a->N, {&a->reb, &a->ebfr, &a->q, &a->P_orb, &a->inc, &a->ecc, &a->argp,
                           &a->mtot, &a->rhost, &a->u1, &a->u2, &a->cfr});
    if (rc) return rc;
    if (a->N == 0) return tri_eval_eb_dev(a, out, nullptr);
    CU(cudaSetDevice(g.device));
    const int64_t N = a->N;
    cudaStream_t s = g.stream;
    g.staging.reset();
    rc = g.staging.reserve((size_t)N * 8 * 16 + (size_t)N * 3 + 8192
                           + (size_t)std::max<int64_t>(out[0].top_cap, 0) * 16
                           + (size_t)std::max<int64_t>(out[1].top_cap, 0) * 16);
    if (rc) return rc;
    tri_eb_args d = *a;
    Col c;
    STAGE(reb) STAGE(ebfr) STAGE(q) STAGE(P_orb) STAGE(inc) STAGE(ecc) STAGE(argp) STAGE(mtot)
    STAGE(rhost) STAGE(u1) STAGE(u2) STAGE(cfr) STAGE(lnprior)
#undef STAGE
    if (a->extra_mask) {
        uint8_t* m = g.staging.take<uint8_t>(N);
        CU(cudaMemcpyAsync(m, a->extra_mask, (size_t)N, cudaMemcpyHostToDevice, s));
        d.extra_mask = m;
    }
    tri_result r[2] = {out[0], out[1]};
    double* h_lnl[2] = {out[0].lnL_out, out[1].lnL_out};
    uint8_t* h_mask[2] = {out[0].mask_out, out[1].mask_out};
    int64_t* h_tidx[2] = {out[0].top_idx, out[1].top_idx};
    double* h_tval[2] = {out[0].top_lnL, out[1].top_lnL};
    bool top[2];
    for (int b = 0; b < 2; ++b) {
        top[b] = out[b].top_cap > 0 && h_tidx[b] && h_tval[b];
        r[b].lnL_out = h_lnl[b] ? g.staging.take<double>(N) : nullptr;
        r[b].mask_out = h_mask[b] ? g.staging.take<uint8_t>(N) : nullptr;
        r[b].top_cap = top[b] ? out[b].top_cap : 0;
        r[b].top_idx = top[b] ? g.staging.take<int64_t>(out[b].top_cap) : nullptr;
        r[b].top_lnL = top[b] ? g.staging.take<double>(out[b].top_cap) : nullptr;
    }
    rc = eval_eb_device(d, r, s);
    if (rc) return rc;
    for (int b = 0; b < 2; ++b) {
        if (h_lnl[b])
            CU(cudaMemcpyAsync(h_lnl[b], r[b].lnL_out, (size_t)N * 8, cudaMemcpyDeviceToHost, s));
        if (h_mask[b])
            CU(cudaMemcpyAsync(h_mask[b], r[b].mask_out, (size_t)N, cudaMemcpyDeviceToHost, s));
        if (top[b] && r[b].n_top > 0) {
            CU(cudaMemcpyAsync(h_tidx[b], r[b].top_idx, (size_t)r[b].n_top * 8,
                               cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(h_tval[b], r[b].top_lnL, (size_t)r[b].n_top * 8,
                               cudaMemcpyDeviceToHost, s));
        }
    }
    CU(cudaStreamSynchronize(s));
    for (int b = 0; b < 2; ++b) {
        if (top[b]) sort_candidates(r[b].n_top, h_tidx[b], h_tval[b]);
        r[b].lnL_out = h_lnl[b];
        r[b].mask_out = h_mask[b];
        r[b].top_cap = out[b].top_cap;
        r[b].top_idx = h_tidx[b];
        r[b].top_lnL = h_tval[b];
        out[b] = r[b];
    }
    return TRI_OK;
}

static int lnl_seam(int eb, int64_t n, const double* body, const double* ebfr, const double* P,
                    const double* inc, const double* a, const double* R_s, const double* u1,
                    const double* u2, const double* ecc, const double* argp, const double* cfr,
                    int32_t is_host, int32_t twin, double* out, double* model_out = nullptr,
                    double* secdepth_out = nullptr, int scalar_rule = 0) {
    int rc = need_ready(true);
    if (rc) return rc;
    if (n < 0) return fail(TRI_EINVAL, "n must be >= 0");
    if (n == 0) return TRI_OK;
    if (!body || !P || !inc || !a || !R_s || !u1 || !u2 || !ecc || !argp || !cfr ||
        (!out && !model_out) || (eb && !ebfr))
        return fail(TRI_EINVAL, "NULL array");
    const size_t npts = (size_t)g.lc.npts;
    if (model_out && (size_t)n * npts > ((size_t)1 << 28))
        return fail(TRI_EINVAL, "model matrix too large (n * npts > 2^28): simulate fewer draws");
    CU(cudaSetDevice(g.device));
    cudaStream_t s = g.stream;
    g.staging.reset();
    rc = g.staging.reserve((size_t)n * 8 * 14 + 8192 + (model_out ? (size_t)n * npts * 8 : 0));
    if (rc) return rc;
    LnlArgs A{};
    A.lc = g.lc; A.tab = g.tab; A.eb = eb; A.companion_is_host = is_host; A.raw = 1;
    A.twin_uniform = twin ? 1 : 0;
    Col c;
#define UP(dst, src)                                           \
    rc = stage(tri_col{src, 1}, n, c, s);                      \
    if (rc) return rc;                                         \
    A.dst = c;
    UP(body, body)
    if (eb) { UP(ebfr, ebfr) }
    UP(P, P) UP(inc, inc) UP(a, a) UP(rhost, R_s) UP(u1, u1) UP(u2, u2) UP(ecc, ecc)
    UP(argp, argp) UP(cfr, cfr)
#undef UP
    double* d_out = g.staging.take<double>(n);
    double* d_model = model_out ? g.staging.take<double>((size_t)n * npts) : nullptr;
    double* d_sec = secdepth_out ? g.staging.take<double>(n) : nullptr;
    g.launches = 0;
    CU(cudaMemsetAsync(g.d_counters, 0, 8 * sizeof(unsigned long long), s));
    A.items = nullptr; A.count = n; A.count_dev = nullptr; A.next = g.d_counters + 2;
    A.out = d_out; A.out_twin = nullptr; A.counters = g.d_counters + 4;
    A.model_out = d_model; A.secdepth_out = d_sec; A.perm = g.d_perm;
    A.scalar_rule = scalar_rule;
    CU(cudaEventRecord(g.ev[0], s));
    CU(cudaEventRecord(g.ev[1], s));
    rc = launch_lnl(A, s);
    if (rc) return rc;
    CU(cudaEventRecord(g.ev[2], s));
    CU(cudaEventRecord(g.ev[3], s));
    if (out) CU(cudaMemcpyAsync(out, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    if (model_out)
        CU(cudaMemcpyAsync(model_out, d_model, (size_t)n * npts * 8, cudaMemcpyDeviceToHost, s));
    if (secdepth_out)
        CU(cudaMemcpyAsync(secdepth_out, d_sec, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    g.timing_valid = true;
    return TRI_OK;
}

int tri_lnl_tp(int64_t n, const double* R_p, const double* P_orb, const double* inc,
               const double* a, const double* R_s, const double* u1, const double* u2,
               const double* ecc, const double* argp, const double* cfr, int32_t is_host,
               double* out) {
    return lnl_seam(0, n, R_p, nullptr, P_orb, inc, a, R_s, u1, u2, ecc, argp, cfr, is_host, 0,
                    out);
}

int tri_lnl_eb(int64_t n, const double* R_EB, const double* EB_fluxratio, const double* P_orb,
               const double* inc, const double* a, const double* R_s, const double* u1,
               const double* u2, const double* ecc, const double* argp, const double* cfr,
               int32_t is_host, int32_t twin, double* out) {
    return lnl_seam(1, n, R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc, argp, cfr,
                    is_host, twin, out);
}

int tri_simulate_tp(int64_t n, const double* R_p, const double* P_orb, const double* inc,
                    const double* a, const double* R_s, const double* u1, const double* u2,
                    const double* ecc, const double* argp, const double* cfr, int32_t is_host,
                    double* flux_out) {
    if (!flux_out) return fail(TRI_EINVAL, "flux_out is NULL");
    return lnl_seam(0, n, R_p, nullptr, P_orb, inc, a, R_s, u1, u2, ecc, argp, cfr, is_host, 1,
                    nullptr, flux_out, nullptr);
}

int tri_simulate_eb(int64_t n, const double* R_EB, const double* EB_fluxratio,
                    const double* P_orb, const double* inc, const double* a, const double* R_s,
                    const double* u1, const double* u2, const double* ecc, const double* argp,
                    const double* cfr, int32_t is_host, int32_t scalar_rule, double* flux_out,
                    double* secdepth_out) {
    if (!flux_out) return fail(TRI_EINVAL, "flux_out is NULL");
    // twin = 1: the secondary-depth cut belongs to lnL_EB_p, not to the simulation
    return lnl_seam(1, n, R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc, argp, cfr,
                    is_host, 1, nullptr, flux_out, secdepth_out, scalar_rule ? 1 : 0);
}

int tri_fetch_lnl(int32_t branch, double* out, int64_t N) {
    int rc = need_ready(false);
    if (rc) return rc;
    if (branch < 0 || branch > 1 || !out) return fail(TRI_EINVAL, "bad argument");
    if (!g.last_lnl[branch] || N != g.last_N)
        return fail(TRI_ESTATE, "no lnL array of that size from the last tri_eval_* call");
    CU(cudaSetDevice(g.device));
    CU(cudaMemcpy(out, g.last_lnl[branch], (size_t)N * 8, cudaMemcpyDeviceToHost));
    return TRI_OK;
}

int tri_log_mean_exp(const double* logw, int64_t n, tri_result* out) {
    int rc = need_ready(false);
    if (rc) return rc;
    if (!out || n < 0 || (n > 0 && !logw)) return fail(TRI_EINVAL, "bad argument");
    LsePartial z{-INFINITY, 0.0, 0, 0};
    if (n == 0) {
        finish_result(z, 0, out);
        return TRI_OK;
    }
    CU(cudaSetDevice(g.device));
    cudaStream_t s = g.stream;
    g.staging.reset();
    rc = g.staging.reserve((size_t)n * 8 + (size_t)lse_blocks(n) * sizeof(LsePartial) + 8192);
    if (rc) return rc;
    double* d = g.staging.take<double>(n);
    LsePartial* parts = g.staging.take<LsePartial>(lse_blocks(n));
    CU(cudaMemcpyAsync(d, logw, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    g.launches = 0;
    rc = launch_lse(d, Col{nullptr, 0}, n, parts, g.d_lse_out, s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(g.h_lse_out, g.d_lse_out, sizeof(LsePartial), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    finish_result(g.h_lse_out[0], n, out);
    return TRI_OK;
}

int tri_last_timing(double* geometry_ms, double* lnl_ms, double* lse_ms, int32_t* launches) {
    int rc = need_ready(false);
    if (rc) return rc;
    if (!g.timing_valid) return fail(TRI_ESTATE, "no timed call yet");
    float a = 0, b = 0, c = 0;
    CU(cudaEventElapsedTime(&a, g.ev[0], g.ev[1]));
    CU(cudaEventElapsedTime(&b, g.ev[1], g.ev[2]));
    CU(cudaEventElapsedTime(&c, g.ev[2], g.ev[3]));
    if (geometry_ms) *geometry_ms = a;
    if (lnl_ms) *lnl_ms = b;
    if (lse_ms) *lse_ms = c;
    if (launches) *launches = g.launches;
    return TRI_OK;
}

int tri_fp64_peak(double* dfma_per_s) {
    int rc = need_ready(false);
    if (rc) return rc;
    if (!dfma_per_s) return fail(TRI_EINVAL, "NULL argument");
    CU(cudaSetDevice(g.device));
    const int threads = 256, blocks = g.sm_count * 8, iters = 1 << 15;
    double* d = nullptr;
    CU(cudaMalloc(&d, (size_t)threads * blocks * sizeof(double)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e0, g.stream));
        dfma_peak_kernel<<<blocks, threads, 0, g.stream>>>(d, iters);
        CU(cudaEventRecord(e1, g.stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        double rate = (double)threads * blocks * 8.0 * iters / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *dfma_per_s = best;
    return TRI_OK;
}

}  // extern "C"
