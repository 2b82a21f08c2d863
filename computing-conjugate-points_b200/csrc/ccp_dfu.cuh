// ccp_dfu.cuh -- localManifold::boundDFU on the GPU (localManifold.cpp:347-409).
//
// One CTA per descriptor (= one call of boundDFU).  Thread t owns column j = t % 2n of the derivative matrix and walks over every
// P-th sub-box (P = blockDim / 2n): it forms u = p + A w over the sub-box, the tridiagonal M(u) in centred form, column j of
// M A[:n,:] and of Ainv Df A, and keeps the running hull of its column in registers; the P partial hulls of a column meet in shared
// memory at the end (min / max: independent of the order, hence bit-identical to the serial oracle, oracle/dfu.hpp).
// Interval arithmetic with static rounding modes as everywhere (ccp_interval.cuh); FP64 vector pipe, no memory traffic to speak of.
#pragma once
#include "ccp_interval.cuh"
#include "../../include/ccp.h"

namespace ccp {

constexpr int DFU_NT = 256;

__device__ __forceinline__ void dfu_ipfma(ival& a, const ival& x, double p) {
    const double l = p >= 0 ? x.lo : x.hi, h = p >= 0 ? x.hi : x.lo;
    a.lo = __fma_rd(l, p, a.lo); a.hi = __fma_ru(h, p, a.hi);
}
__device__ __forceinline__ ival dfu_sqr(const ival& a) {
    if (a.lo >= 0) return mk(__dmul_rd(a.lo, a.lo), __dmul_ru(a.hi, a.hi));
    if (a.hi <= 0) return mk(__dmul_rd(a.hi, a.hi), __dmul_ru(a.lo, a.lo));
    return mk(0.0, dmax(__dmul_ru(a.lo, a.lo), __dmul_ru(a.hi, a.hi)));
}
__device__ __forceinline__ ival dfu_divn(const ival& a, int n) { return mk(__ddiv_rd(a.lo, (double)n), __ddiv_ru(a.hi, (double)n)); }

template <int N>
struct DfuSmem {
    static constexpr int D = 2 * N;
    ival b[N], c[N], p[D], U[D], Ainv[D][D], K0[D][D], m0[N], b3[N], c2[N];
    double A[D][D];
    ival parts[N][CCP_DFU_MAX_SUBDIV];
    ival red[DFU_NT / D][D][D];        // partial hulls: [group][row][column]
};

template <int N>
__global__ void __launch_bounds__(DFU_NT) dfu_kernel(const ccp_dfu_desc* __restrict__ descs, ccp_dfu_result* __restrict__ results, int count) {
    constexpr int D = 2 * N, P = DFU_NT / D;
    extern __shared__ __align__(16) unsigned char dfu_raw[];
    DfuSmem<N>& S = *reinterpret_cast<DfuSmem<N>*>(dfu_raw);
    const int tid = threadIdx.x;
    for (int q = blockIdx.x; q < count; q += gridDim.x) {
        const ccp_dfu_desc& d = descs[q];
        ccp_dfu_result& r = results[q];
        if (d.n != N) continue;                                   // another instantiation handles it
        const int Ns = d.subdivision;
        if (Ns < 1 || Ns > CCP_DFU_MAX_SUBDIV) { if (tid == 0) { r.status = CCP_ST_BAD_DESC; r.parts = 0; } continue; }
        __syncthreads();
        // ---- stage the descriptor ----
        if (tid < N) { S.b[tid] = mk(d.b[tid].lo, d.b[tid].hi); S.c[tid] = tid < N - 1 ? mk(d.c[tid].lo, d.c[tid].hi) : pt(0.0); }
        if (tid < D) { S.p[tid] = mk(d.p[tid].lo, d.p[tid].hi); S.U[tid] = mk(d.U[tid].lo, d.U[tid].hi); }
        for (int t = tid; t < D * D; t += DFU_NT) {
            const int i = t / D, j = t % D;
            S.A[i][j] = d.A[i * CCP_MAX_DIM + j];
            S.Ainv[i][j] = mk(d.Ainv[i * CCP_MAX_DIM + j].lo, d.Ainv[i * CCP_MAX_DIM + j].hi);
        }
        __syncthreads();
        const int idx0 = d.stable ? N : 0;
        // ---- parts of the subdivided coordinates (utils.cpp:3-10), constants of M, constant part of the product ----
        for (int t = tid; t < N * Ns; t += DFU_NT) {
            const int cq = t / Ns, k = t % Ns;
            const ival x = S.U[idx0 + cq], xl = pt(x.lo), xr = pt(x.hi);
            S.parts[cq][k] = iadd(iadd(xl, dfu_divn(imulk(isub(xr, xl), k), Ns)), dfu_divn(isub(x, xl), Ns));
        }
        if (tid >= 64 && tid < 64 + N) {
            const int i = tid - 64;
            ival s = S.b[i];
            if (i > 0) s = iadd(s, S.c[i - 1]);
            if (i < N - 1) s = iadd(s, S.c[i]);
            S.m0[i] = mk(0.25 * s.lo, 0.25 * s.hi);
            S.b3[i] = imulk(S.b[i], 3);
            S.c2[i] = i < N - 1 ? imulk(S.c[i], 2) : pt(0.0);
        }
        for (int t = tid; t < D * D; t += DFU_NT) {
            const int i = t / D, j = t % D;
            ival a = pt(0.0);
            for (int k = 0; k < N; k++) dfu_ipfma(a, S.Ainv[i][k], S.A[N + k][j]);
            S.K0[i][j] = a;
        }
        __syncthreads();
        long total = 1;
        for (int cq = 0; cq < N; cq++) total *= Ns;
        // ---- sub-boxes ----
        const int j = tid % D, grp = tid / D;
        ival H[D];
        bool have = false;
        if (grp < P) {
            for (long pi = grp; pi < total; pi += P) {
                ival Up[D];
#pragma unroll
                for (int i = 0; i < D; i++) Up[i] = S.U[i];
                long rest = pi;
#pragma unroll
                for (int cq = 0; cq < N; cq++) { const ival pv = S.parts[cq][rest % Ns]; rest /= Ns;
#pragma unroll
                    for (int i = 0; i < D; i++) if (i == idx0 + cq) Up[i] = pv; }
                ival w[N], w2[N];
#pragma unroll
                for (int k = 0; k < N; k++) {
                    ival a = pt(0.0);
#pragma unroll
                    for (int m = 0; m < D; m++) dfu_ipfma(a, Up[m], S.A[k][m]);
                    w[k] = isub(iadd(S.p[k], a), pt(0.5));
                    w2[k] = dfu_sqr(w[k]);
                }
                ival Md[N], Mo[N];
#pragma unroll
                for (int i = 0; i < N; i++) {
                    ival a = imul(S.b3[i], w2[i]);
                    if (i > 0)     a = iadd(a, imul(S.c[i - 1], w2[i - 1]));
                    if (i < N - 1) a = iadd(a, imul(S.c[i], w2[i + 1 < N ? i + 1 : i]));
                    Md[i] = isub(a, S.m0[i]);
                    Mo[i] = i < N - 1 ? imul(S.c2[i], imul(w[i], w[i + 1 < N ? i + 1 : i])) : pt(0.0);
                }
                ival MR[N];
#pragma unroll
                for (int k = 0; k < N; k++) {
                    ival a = pt(0.0);
                    if (k > 0)     dfu_ipfma(a, Mo[k > 0 ? k - 1 : 0], S.A[k > 0 ? k - 1 : 0][j]);
                    dfu_ipfma(a, Md[k], S.A[k][j]);
                    if (k < N - 1) dfu_ipfma(a, Mo[k], S.A[k + 1 < N ? k + 1 : k][j]);
                    MR[k] = a;
                }
#pragma unroll
                for (int i = 0; i < D; i++) {
                    ival a = S.K0[i][j];
#pragma unroll
                    for (int k = 0; k < N; k++) a = iadd(a, imul(S.Ainv[i][N + k], MR[k]));
                    H[i] = have ? ihull(H[i], a) : a;
                }
                have = true;
            }
#pragma unroll
            for (int i = 0; i < D; i++) S.red[grp][i][j] = have ? H[i] : mk(__longlong_as_double(0x7ff0000000000000LL), __longlong_as_double(0xfff0000000000000LL));   // empty: [+inf, -inf]
        }
        __syncthreads();
        for (int t = tid; t < D * D; t += DFU_NT) {
            const int i = t / D, jj = t % D;
            ival a = S.red[0][i][jj];
            for (int g2 = 1; g2 < P; g2++) a = ihull(a, S.red[g2][i][jj]);
            r.DF[i * CCP_MAX_DIM + jj].lo = a.lo; r.DF[i * CCP_MAX_DIM + jj].hi = a.hi;
        }
        if (tid == 0) { r.status = CCP_OK; r.parts = (int)total; }
    }
}

} // namespace ccp
