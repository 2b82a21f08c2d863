// oracle/dfu.hpp -- TEST INFRASTRUCTURE (CPU oracle).  Not part of the product path.
//
// Restatement of localManifold::boundDFU (localManifold.cpp:347-409): the interval hull, over the N^n sub-boxes of the
// neighbourhood U (part(), utils.cpp:3-10; getSubdivision, utils.cpp:309-321), of the derivative of the local vector field
//     localVField::derivative(w) = Ainv * Df(p + A w) * A                                     (localVField.h:31-35)
// where f is the 2n-dimensional front field (ODE_functions.cpp:206,239 without the variational part), so that
//     Df(x) = [0 I; M(u) 0],   M_ii = 3 b_i w_i^2 + sum_{j~i} c_ij w_j^2 - (b_i + sum_j c_ij)/4,   M_ij = 2 c_ij w_i w_j,   w = u - 1/2
// (the centred form of SURVEY.md B.1: every w occurs once per entry, so the interval evaluation is the exact range; CAPD differentiates
// the formula string by automatic differentiation [CAPD-recollection], which encloses the same matrix).  Same operations in the same
// order as the CUDA kernel csrc/ccp_dfu.cuh; the hull makes the result independent of the order of the sub-boxes.
#pragma once
#include "ival.hpp"
#include "../include/ccp.h"

namespace orc {
namespace dfu {

static inline void ipfma(ival& a, const ival& x, double p) {
    const double l = p >= 0 ? x.lo : x.hi, h = p >= 0 ? x.hi : x.lo;
    a.lo = fma_rd(l, p, a.lo); a.hi = fma_ru(h, p, a.hi);
}
static inline ival mulk(const ival& a, int k) { return ival(mul_rd((double)k, a.lo), mul_ru((double)k, a.hi)); }
static inline ival divn(const ival& a, int n) { return ival(div_rd(a.lo, (double)n), div_ru(a.hi, (double)n)); }
// part(x, k, N) = x.left() + k*(x.right() - x.left())/N + (x - x.left())/N                      (utils.cpp:3-10)
static inline ival part(const ival& x, int k, int N) {
    const ival xl(x.lo), xr(x.hi);
    return (xl + divn(mulk(xr - xl, k), N)) + divn(x - xl, N);
}

static inline void run(const ccp_dfu_desc& d, ccp_dfu_result& res) {
    std::memset(&res, 0, sizeof(res));
    const int n = d.n, D = 2 * n, N = d.subdivision;
    if (n < 2 || n > CCP_MAX_N || N < 1 || N > CCP_DFU_MAX_SUBDIV) { res.status = CCP_ST_BAD_DESC; return; }
    ival b[CCP_MAX_N], c[CCP_MAX_N], p[CCP_MAX_DIM], U[CCP_MAX_DIM], Ainv[CCP_MAX_DIM][CCP_MAX_DIM];
    double A[CCP_MAX_DIM][CCP_MAX_DIM];
    for (int i = 0; i < n; i++) b[i] = ival(d.b[i].lo, d.b[i].hi);
    for (int i = 0; i < n - 1; i++) c[i] = ival(d.c[i].lo, d.c[i].hi);
    for (int i = 0; i < D; i++) {
        p[i] = ival(d.p[i].lo, d.p[i].hi); U[i] = ival(d.U[i].lo, d.U[i].hi);
        for (int j = 0; j < D; j++) { A[i][j] = d.A[i * CCP_MAX_DIM + j]; Ainv[i][j] = ival(d.Ainv[i * CCP_MAX_DIM + j].lo, d.Ainv[i * CCP_MAX_DIM + j].hi); }
    }
    const int idx0 = d.stable ? n : 0;
    // parts of the subdivided coordinates; constant parts of M and of the product
    ival parts[CCP_MAX_N][CCP_DFU_MAX_SUBDIV], m0[CCP_MAX_N], b3[CCP_MAX_N], c2[CCP_MAX_N], K0[CCP_MAX_DIM][CCP_MAX_DIM];
    for (int q = 0; q < n; q++) for (int k = 0; k < N; k++) parts[q][k] = part(U[idx0 + q], k, N);
    for (int i = 0; i < n; i++) {
        ival s = b[i];
        if (i > 0) s = s + c[i - 1];
        if (i < n - 1) s = s + c[i];
        m0[i] = ival(0.25 * s.lo, 0.25 * s.hi);                       // (b_i + sum c)/4
        b3[i] = mulk(b[i], 3);
        if (i < n - 1) c2[i] = mulk(c[i], 2);
    }
    for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) { ival a(0.0); for (int k = 0; k < n; k++) ipfma(a, Ainv[i][k], A[n + k][j]); K0[i][j] = a; }
    long total = 1; for (int q = 0; q < n; q++) total *= N;
    ival H[CCP_MAX_DIM][CCP_MAX_DIM];
    for (long pi = 0; pi < total; pi++) {
        ival Up[CCP_MAX_DIM];
        for (int i = 0; i < D; i++) Up[i] = U[i];
        long rest = pi;
        for (int q = 0; q < n; q++) { Up[idx0 + q] = parts[q][rest % N]; rest /= N; }          // part_list[q] = (j / N^q) % N
        ival w[CCP_MAX_N], w2[CCP_MAX_N];
        for (int k = 0; k < n; k++) {
            ival a(0.0);
            for (int m = 0; m < D; m++) ipfma(a, Up[m], A[k][m]);
            w[k] = (p[k] + a) - ival(0.5);
            w2[k] = sqr(w[k]);
        }
        ival Md[CCP_MAX_N], Mo[CCP_MAX_N];
        for (int i = 0; i < n; i++) {
            ival a = b3[i] * w2[i];
            if (i > 0)     a = a + c[i - 1] * w2[i - 1];
            if (i < n - 1) a = a + c[i] * w2[i + 1];
            Md[i] = a - m0[i];
            if (i < n - 1) Mo[i] = c2[i] * (w[i] * w[i + 1]);
        }
        for (int j = 0; j < D; j++) {
            ival MR[CCP_MAX_N];
            for (int k = 0; k < n; k++) {
                ival a(0.0);
                if (k > 0)     ipfma(a, Mo[k - 1], A[k - 1][j]);
                ipfma(a, Md[k], A[k][j]);
                if (k < n - 1) ipfma(a, Mo[k], A[k + 1][j]);
                MR[k] = a;
            }
            for (int i = 0; i < D; i++) {
                ival a = K0[i][j];
                for (int k = 0; k < n; k++) a = a + Ainv[i][n + k] * MR[k];
                H[i][j] = pi == 0 ? a : hull(H[i][j], a);
            }
        }
    }
    res.status = CCP_OK; res.parts = (int32_t)total;
    for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) { res.DF[i * CCP_MAX_DIM + j].lo = H[i][j].lo; res.DF[i * CCP_MAX_DIM + j].hi = H[i][j].hi; }
}

} // namespace dfu
} // namespace orc
