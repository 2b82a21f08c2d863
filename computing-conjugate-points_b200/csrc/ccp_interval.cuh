// ccp_interval.cuh -- device interval arithmetic for sm_100a.
//
// Every bound is produced by ONE instruction with a static rounding mode (DADD.RM/.RP, DMUL.RM/.RP,
// DFMA.RM/.RP in SASS via __dadd_rd/_ru, __dmul_rd/_ru, __fma_rd/_ru): no rounding-mode switches, which
// is the architectural advantage over the CAPD/x87-SSE path of the reference.
// Two number formats:
//   ival  [lo,hi]                      inf-sup, used for linear algebra, Horner and determinants
//   mt    (m,t), t >= |m| + radius     used inside Cauchy products: 4 DFMA per term
//         sum_j x_j*y_j  in  [S_lo - R, S_hi + R],  S_lo/S_hi = RD/RU chains of m_x*m_y,
//         R = RU(sum t_x*t_y) - RD(sum |m_x||m_y|)   (>= sum of |m_x| r_y + r_x |m_y| + r_x r_y)
// The op sequences are mirrored one-for-one by the CPU oracle (oracle/c1.hpp); the GPU/CPU
// results are compared bit for bit in tests/test_gpu_parity.py.
#pragma once
#include <cuda_runtime.h>

namespace ccp {

struct __align__(16) ival { double lo, hi; };     // 16-byte aligned: one LDS.128 / STS.128 per interval
struct __align__(16) mt   { double m, t; };

__device__ __forceinline__ double dmin(double a, double b) { return (b < a) ? b : a; }   // same tie/zero behaviour as std::min
__device__ __forceinline__ double dmax(double a, double b) { return (a < b) ? b : a; }   // same as std::max

__device__ __forceinline__ ival mk(double lo, double hi) { ival r; r.lo = lo; r.hi = hi; return r; }
__device__ __forceinline__ ival pt(double x) { return mk(x, x); }
__device__ __forceinline__ ival sym(double r) { return mk(-r, r); }

__device__ __forceinline__ ival iadd(const ival& a, const ival& b) { return mk(__dadd_rd(a.lo, b.lo), __dadd_ru(a.hi, b.hi)); }
__device__ __forceinline__ ival isub(const ival& a, const ival& b) { return mk(__dadd_rd(a.lo, -b.hi), __dadd_ru(a.hi, -b.lo)); }
__device__ __forceinline__ ival ineg(const ival& a) { return mk(-a.hi, -a.lo); }
__device__ __forceinline__ ival imul(const ival& a, const ival& b) {
    double l = dmin(dmin(__dmul_rd(a.lo, b.lo), __dmul_rd(a.lo, b.hi)), dmin(__dmul_rd(a.hi, b.lo), __dmul_rd(a.hi, b.hi)));
    double h = dmax(dmax(__dmul_ru(a.lo, b.lo), __dmul_ru(a.lo, b.hi)), dmax(__dmul_ru(a.hi, b.lo), __dmul_ru(a.hi, b.hi)));
    return mk(l, h);
}
// out-of-line copy for the many cold call sites (determinants, set update): keeps the step loop's code small enough
// for the instruction cache; the arithmetic is identical
__device__ __noinline__ ival imul_nl(const ival& a, const ival& b) { return imul(a, b); }
__device__ __forceinline__ ival ihull(const ival& a, const ival& b) { return mk(dmin(a.lo, b.lo), dmax(a.hi, b.hi)); }
__device__ __forceinline__ double imag(const ival& a) { return dmax(fabs(a.lo), fabs(a.hi)); }
__device__ __forceinline__ bool idefinite(const ival& a) { return a.lo > 0 || a.hi < 0; }
__device__ __forceinline__ bool isubset(const ival& a, const ival& b) { return b.lo <= a.lo && a.hi <= b.hi; }
__device__ __forceinline__ bool ifinite(const ival& a) { return isfinite(a.lo) && isfinite(a.hi); }
__device__ __forceinline__ ival idivk(const ival& a, int k) { return mk(__ddiv_rd(a.lo, (double)k), __ddiv_ru(a.hi, (double)k)); }
__device__ __forceinline__ ival imulk(const ival& a, int k) { return mk(__dmul_rd((double)k, a.lo), __dmul_ru((double)k, a.hi)); }
__device__ __forceinline__ double mid_ru(const ival& a) { return __dadd_ru(0.5 * a.lo, 0.5 * a.hi); }

__device__ __forceinline__ ival mt2iv(const mt& a) { double r = __dadd_ru(a.t, -fabs(a.m)); return mk(__dadd_rd(a.m, -r), __dadd_ru(a.m, r)); }
// m = RU(lo/2 + hi/2) is never below the exact midpoint (hi/2 is exact, or rounded up for subnormals; lo/2 enters the FMA unrounded), hence
// m - lo >= hi - m and r = RU(m - lo) alone bounds both distances: no comparison on the critical chains.  (For normal numbers this is
// bit-identical to max(RU(hi - m), RU(m - lo)): rounding is monotone.)
__device__ __forceinline__ mt iv2mt(const ival& a) {
    mt o; o.m = __fma_ru(0.5, a.lo, __dmul_ru(0.5, a.hi));       // one instruction less than RU(RU(lo/2) + RU(hi/2)); the same bits whenever lo/2 is exact (always, but for subnormals)
    const double r = __dadd_ru(o.m, -a.lo);
    o.t = __dadd_ru(fabs(o.m), r);
    return o;
}

struct acc4 { double slo, shi, T, Q; };
__device__ __forceinline__ acc4 acc_zero() { acc4 a; a.slo = 0; a.shi = 0; a.T = 0; a.Q = 0; return a; }
__device__ __forceinline__ void acc_term(acc4& a, const mt& x, const mt& y) {
    a.slo = __fma_rd(x.m, y.m, a.slo);
    a.shi = __fma_ru(x.m, y.m, a.shi);
    a.T   = __fma_ru(x.t, y.t, a.T);
    a.Q   = __fma_rd(fabs(x.m), fabs(y.m), a.Q);
}
__device__ __forceinline__ void acc_term_neg(acc4& a, const mt& x, const mt& y) {   // += x * (-y)
    a.slo = __fma_rd(x.m, -y.m, a.slo);
    a.shi = __fma_ru(x.m, -y.m, a.shi);
    a.T   = __fma_ru(x.t, y.t, a.T);
    a.Q   = __fma_rd(fabs(x.m), fabs(y.m), a.Q);
}
// R = RU(T - Q) >= 0 always: every term of T (round-up chain of t_x t_y) dominates the matching term of Q (round-down chain of
// |m_x||m_y| <= t_x t_y), so no sign guard is needed; a NaN propagates into both bounds and is caught by the finiteness checks.
__device__ __forceinline__ ival acc_finish(const acc4& a) {
    const double R = __dadd_ru(a.T, -a.Q);
    return mk(__dadd_rd(a.slo, -R), __dadd_ru(a.shi, R));
}

// Horner steps for tau >= 0
__device__ __forceinline__ ival horner_pt(const ival& acc, double tau, const ival& c) { return mk(__fma_rd(acc.lo, tau, c.lo), __fma_ru(acc.hi, tau, c.hi)); }
__device__ __forceinline__ ival horner_rg(const ival& acc, double tl, double th, const ival& c) {
    return mk(__fma_rd(acc.lo, acc.lo >= 0 ? tl : th, c.lo), __fma_ru(acc.hi, acc.hi >= 0 ? th : tl, c.hi));
}

} // namespace ccp
