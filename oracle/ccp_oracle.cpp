// oracle/ccp_oracle.cpp -- TEST INFRASTRUCTURE: C entry points of the CPU oracle (liboracle.so).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library.  It is the checker, never the product: the product path lives in
// computing-conjugate-points_b200/csrc and fails loudly without a GPU.
//
// PARITY PINNING: the reference cannot be built here (needs CAPD 5.0.59, absent) and ships no tests;
// the only reference output that exists is plot_det.txt (15,176 det range enclosures of its n=3
// default run).  tests/test_oracle_golden.py pins both oracle modes against the committed digest of
// that file (tests/golden/plot_det_digest.npz).  Everything CAPD-internal is a mathematical
// restatement => parity is "pinned on det enclosures of config 1, otherwise unpinned".
//
//   mode 0 ("ref"): reference-structured -- n separate d=4n dimensional C^0 Lohner integrations of
//                   f_linearize, curve evaluation on the grid, topFrame series + countZeros.
//   mode 1 ("c1") : structured C^1 algorithm, op-for-op what the CUDA kernel does.
#include "ref.hpp"
#include "c1.hpp"
#include "c2.hpp"
#include "dfu.hpp"
#include "subdiv.hpp"
#include <thread>
#include <atomic>
#include <chrono>

using namespace orc;

extern "C" {

int orc_run_front(const ccp_front_desc* d, int mode, ccp_front_result* res, ccp_series_rec* series, size_t cap, size_t* len, int threads) {
    RoundUpGuard guard;
    std::vector<ccp_series_rec> ser;
    bool want = series != nullptr && cap > 0;
    if (mode == 1) c1::run_front(*d, *res, want ? &ser : nullptr);
    else if (mode == 2) c2::run_front(*d, *res, want ? &ser : nullptr);
    else           ref::run_front(*d, *res, want ? &ser : nullptr, threads);
    if (len) *len = ser.size();
    if (want) { size_t m = std::min(cap, ser.size()); std::memcpy(series, ser.data(), m * sizeof(ccp_series_rec)); }
    return 0;
}

// batch over fronts with host threads; returns wall seconds in *seconds
int orc_run_batch(const ccp_front_desc* d, size_t n_fronts, int mode, ccp_front_result* res, int threads, double* seconds) {
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if ((size_t)nt > n_fronts) nt = (int)std::max<size_t>(1, n_fronts);
    std::atomic<size_t> next(0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        RoundUpGuard guard;                         // rounding mode is per thread
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n_fronts) break;
            if (mode == 1) c1::run_front(d[i], res[i], nullptr);
            else if (mode == 2) c2::run_front(d[i], res[i], nullptr);
            else           ref::run_front(d[i], res[i], nullptr, 1);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

// BASELINE config 4 (twin of ccp_frame_count_subdivided): sub-box descriptors by part(), every sub-box through `mode`, hull / tally.
// `sub_out` (optional, n_fronts * parts^n records) receives the sub-box records.
int orc_subdivided(const ccp_front_desc* d, size_t n_fronts, int parts, int which, int mode, ccp_subdiv_result* res, ccp_front_result* sub_out, int threads) {
    if (n_fronts == 0) return 0;
    int total = 1;
    for (int q = 0; q < d[0].n; q++) total *= parts;
    std::vector<ccp_front_desc> sub((size_t)total * n_fronts);
    std::vector<ccp_front_result> rec((size_t)total * n_fronts);
    {
        RoundUpGuard guard;
        for (size_t f = 0; f < n_fronts; f++) for (int j = 0; j < total; j++) sub[f * total + j] = subdiv::sub_desc(d[f], j, parts, which);
    }
    double sec = 0;
    orc_run_batch(sub.data(), sub.size(), mode, rec.data(), threads, &sec);
    for (size_t f = 0; f < n_fronts; f++) subdiv::reduce(rec.data() + f * total, total, res[f]);
    if (sub_out) std::memcpy(sub_out, rec.data(), rec.size() * sizeof(ccp_front_result));
    return 0;
}

int orc_bound_dfu(const ccp_dfu_desc* d, size_t n, ccp_dfu_result* res) {
    RoundUpGuard guard;
    for (size_t i = 0; i < n; i++) dfu::run(d[i], res[i]);
    return 0;
}

// ---- primitive probes for unit tests (results bit-compared with the CUDA intrinsics and mpmath) ----
// op: 0 add, 1 sub, 2 mul, 3 div, 4 sqr(a), 5 hull
int orc_ival_op(int op, double alo, double ahi, double blo, double bhi, double* out) {
    RoundUpGuard guard;
    ival a(alo, ahi), b(blo, bhi), r;
    switch (op) {
        case 0: r = a + b; break;
        case 1: r = a - b; break;
        case 2: r = a * b; break;
        case 3: r = a / b; break;
        case 4: r = sqr(a); break;
        case 5: r = hull(a, b); break;
        default: return -1;
    }
    out[0] = r.lo; out[1] = r.hi;
    return 0;
}
// mt-form Cauchy dot product used by mode c1 / the kernel: sum_j x_j*y_j, inputs as (m,t) pairs
int orc_mt_dot(const double* xm, const double* xt, const double* ym, const double* yt, int len, double* out) {
    RoundUpGuard guard;
    c1::acc4 a;
    for (int j = 0; j < len; j++) { c1::mt x{xm[j], xt[j]}, y{ym[j], yt[j]}; c1::acc_term(a, x, y); }
    ival r = c1::acc_finish(a);
    out[0] = r.lo; out[1] = r.hi;
    return 0;
}
// same op codes as ccp_debug_primitives (csrc/ccp_engine.cu)
int orc_debug_primitives(int op, const double* in, double* out, size_t n) {
    RoundUpGuard guard;
    using namespace c1;
    for (size_t i = 0; i < n; i++) {
        const double* x = in + 8 * i; double* o = out + 4 * i;
        o[0] = o[1] = o[2] = o[3] = 0.0;
        switch (op) {
            case 0: o[0] = add_rd(x[0], x[1]); break;
            case 1: o[0] = add_ru(x[0], x[1]); break;
            case 2: o[0] = mul_rd(x[0], x[1]); break;
            case 3: o[0] = mul_ru(x[0], x[1]); break;
            case 4: o[0] = fma_rd(x[0], x[1], x[2]); break;
            case 5: o[0] = fma_ru(x[0], x[1], x[2]); break;
            case 6: o[0] = div_rd(x[0], x[1]); break;
            case 7: o[0] = div_ru(x[0], x[1]); break;
            case 10: { ival r = ival(x[0], x[1]) * ival(x[2], x[3]); o[0] = r.lo; o[1] = r.hi; break; }
            case 11: { mt r = iv2mt(ival(x[0], x[1])); o[0] = r.m; o[1] = r.t; break; }
            case 12: { ival r = mt2iv(mt{x[0], x[1]}); o[0] = r.lo; o[1] = r.hi; break; }
            case 13: { acc4 a; acc_term(a, mt{x[0], x[1]}, mt{x[2], x[3]}); acc_term_neg(a, mt{x[4], x[5]}, mt{x[6], x[7]}); ival r = acc_finish(a); o[0] = r.lo; o[1] = r.hi; break; }
            case 14: { ival r = divk(ival(x[0], x[1]), (int)x[2]); o[0] = r.lo; o[1] = r.hi; break; }
            case 15: { ival r = horner_rg(ival(x[0], x[1]), x[2], x[3], ival(x[4], x[5])); o[0] = r.lo; o[1] = r.hi; break; }
            case 16: { ival r = ival(x[0], x[1]) - ival(x[2], x[3]); o[0] = r.lo; o[1] = r.hi; break; }
            case 17: { ival r = ival(x[0], x[1]) + ival(x[2], x[3]); o[0] = r.lo; o[1] = r.hi; break; }
            case 18: { ival r = mulk(ival(x[0], x[1]), (int)x[2]); o[0] = r.lo; o[1] = r.hi; break; }
            case 19: { ival r = horner_pt(ival(x[0], x[1]), x[2], ival(x[4], x[5])); o[0] = r.lo; o[1] = r.hi; break; }
            case 20: { Frame m; for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { int q = (a * 3 + b) % 7; m[a][b] = ival(std::min(x[q], x[q + 1]), std::max(x[q], x[q + 1])); }
                       ival r = detN(m, 3); o[0] = r.lo; o[1] = r.hi; break; }
            case 21: { mt r = scalek(mt{x[0], x[1]}, (int)x[2]); o[0] = r.m; o[1] = r.t; break; }
            case 22: { ival a(x[0], x[1]); c2::ipfma(a, ival(x[2], x[3]), x[4]); o[0] = a.lo; o[1] = a.hi; break; }
            default: return -1;
        }
    }
    return 0;
}
int orc_iv2mt(double lo, double hi, double* out) { RoundUpGuard guard; c1::mt m = c1::iv2mt(ival(lo, hi)); out[0] = m.m; out[1] = m.t; return 0; }
int orc_mt2iv(double m, double t, double* out) { RoundUpGuard guard; ival r = c1::mt2iv(c1::mt{m, t}); out[0] = r.lo; out[1] = r.hi; return 0; }
int orc_det(const double* lo, const double* hi, int n, double* out) {
    RoundUpGuard guard;
    Frame F;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) F[i][j] = ival(lo[i * n + j], hi[i * n + j]);
    ival r = detN(F, n);
    out[0] = r.lo; out[1] = r.hi;
    return 0;
}
int orc_majorant(const ccp_front_desc* d, double* out) {
    RoundUpGuard guard;
    ival b[CCP_MAX_N], c[CCP_MAX_N];
    for (int i = 0; i < d->n; i++) b[i] = ival(d->b[i].lo, d->b[i].hi);
    for (int i = 0; i < d->n - 1; i++) c[i] = ival(d->c[i].lo, d->c[i].hi);
    c1::Majorant m = c1::c1_majorant(d->n, b, c, d->order, std::ldexp(1.0, -d->step_log2));
    out[0] = m.rho_phi; out[1] = m.rho_phi_d; out[2] = m.rho_psi; out[3] = m.rho_psi_d;
    out[4] = m.Gs_u.lo; out[5] = m.Gs_u.hi; out[6] = m.Gs_v.lo; out[7] = m.Gs_v.hi;
    return m.ok ? 0 : 1;
}

} // extern "C"
