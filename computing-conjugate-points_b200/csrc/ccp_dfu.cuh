// ccp_dfu.cuh -- localManifold::boundDFU on the GPU (localManifold.cpp:347-409).
//
// One CTA per descriptor (= one call of boundDFU); a group of 2n lanes owns one sub-box at a time and shares u = p + A w and the
// tridiagonal M(u) (centred form) of the sub-box through warp shuffles; each lane keeps the running hull of one column of Ainv Df A
// in registers (details at dfu_kernel).  Products with parameter / Ainv entries use their sign class (2 multiplications, same end points).
// Interval arithmetic with static rounding modes as everywhere (ccp_interval.cuh); FP64 vector pipe, no memory traffic to speak of.
#pragma once
#include "ccp_interval.cuh"
#include "../../include/ccp.h"

namespace ccp {

constexpr int DFU_NT = 256;
#ifndef CCP_DFU_MINB
#define CCP_DFU_MINB 2
#endif
constexpr int DFU_MINB = CCP_DFU_MINB;      // resident CTAs per SM the register allocation aims at

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

// interval with its sign class: the product with it takes 2 directed multiplications instead of 8 and has the same end points as
// the general product (rounding is monotone), ccp_kernel.cuh:imul_par
struct DfuSPar { double pl, ph; int neg, straddle; ival c; };
__device__ __forceinline__ DfuSPar dfu_spar(const ival& c) {
    DfuSPar s; s.c = c; s.straddle = (c.lo < 0 && c.hi > 0) ? 1 : 0; s.neg = (c.hi <= 0 && c.lo < 0) ? 1 : 0;
    s.pl = s.neg ? -c.hi : c.lo; s.ph = s.neg ? -c.lo : c.hi;
    return s;
}
__device__ __forceinline__ ival dfu_mul_sp(const DfuSPar& s, const ival& a) {      // = imul(s.c, a)
    if (s.straddle) return imul(s.c, a);
    const double lo = __dmul_rd(a.lo, a.lo >= 0 ? s.pl : s.ph), hi = __dmul_ru(a.hi, a.hi >= 0 ? s.ph : s.pl);
    return s.neg ? mk(-hi, -lo) : mk(lo, hi);
}
__device__ __forceinline__ ival dfu_shfl(const ival& a, int src) {
    return mk(__shfl_sync(0xffffffffu, a.lo, src), __shfl_sync(0xffffffffu, a.hi, src));
}

template <int N>
struct DfuSmem {
    static constexpr int D = 2 * N;
    static constexpr int G = 32 / D;                     // sub-boxes per warp at a time: a group of D lanes (one per column) owns one sub-box
    static constexpr int P = (DFU_NT / 32) * G;
    ival b[N], c[N], p[D], U[D], Ainv[D][D], K0[D][D], m0[N];
    DfuSPar sb3[N], sc[N], sc2[N], sAinv[D][N];          // 3 b_i, c_i, 2 c_i, Ainv[i][N + k] with their sign classes
    double A[D][D];
    ival parts[N][CCP_DFU_MAX_SUBDIV];
    ival red[P][D][D];                                   // partial hulls: [group][row][column]
};

// One CTA per descriptor (= one call of boundDFU).  A group of D = 2n lanes owns one sub-box at a time: lane k < n forms w_k = u_k - 1/2
// over the sub-box, w_k^2 and row k of the tridiagonal M(u) (neighbours' w through shuffles); the rows are broadcast inside the group and
// lane j then forms column j of M A[:n,:] and of Ainv Df A and keeps the running hull of its column in registers.  The partial hulls of
// the groups meet in shared memory at the end (min / max: independent of the order, hence bit-identical to the serial oracle).
template <int N>
__global__ void __launch_bounds__(DFU_NT, DFU_MINB) dfu_kernel(const ccp_dfu_desc* __restrict__ descs, ccp_dfu_result* __restrict__ results, int count) {
    constexpr int D = 2 * N, G = DfuSmem<N>::G, P = DfuSmem<N>::P;
    extern __shared__ __align__(16) unsigned char dfu_raw[];
    DfuSmem<N>& S = *reinterpret_cast<DfuSmem<N>*>(dfu_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gl = lane / D < G ? lane / D : G - 1;      // group of this lane; the 32 - G*D spare lanes shadow the last group
    const bool act = lane < G * D;
    const int lbase = gl * D;                            // first lane of the group
    const int j = act ? lane - lbase : D - 1;            // column of this lane
    const int k = j < N ? j : N - 1;                     // row of M this lane forms (lanes >= n repeat the last row; never read)
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
            const int i = t / D, jj = t % D;
            S.A[i][jj] = d.A[i * CCP_MAX_DIM + jj];
            S.Ainv[i][jj] = mk(d.Ainv[i * CCP_MAX_DIM + jj].lo, d.Ainv[i * CCP_MAX_DIM + jj].hi);
        }
        __syncthreads();
        const int idx0 = d.stable ? N : 0;
        // ---- parts of the subdivided coordinates (utils.cpp:3-10), constants of M, constant part of the product ----
        for (int t = tid; t < N * Ns; t += DFU_NT) {
            const int cq = t / Ns, kk = t % Ns;
            const ival x = S.U[idx0 + cq], xl = pt(x.lo), xr = pt(x.hi);
            S.parts[cq][kk] = iadd(iadd(xl, dfu_divn(imulk(isub(xr, xl), kk), Ns)), dfu_divn(isub(x, xl), Ns));
        }
        if (tid >= 64 && tid < 64 + N) {
            const int i = tid - 64;
            ival s = S.b[i];
            if (i > 0) s = iadd(s, S.c[i - 1]);
            if (i < N - 1) s = iadd(s, S.c[i]);
            S.m0[i] = mk(0.25 * s.lo, 0.25 * s.hi);
            S.sb3[i] = dfu_spar(imulk(S.b[i], 3));
            S.sc[i] = dfu_spar(S.c[i]);
            S.sc2[i] = dfu_spar(i < N - 1 ? imulk(S.c[i], 2) : pt(0.0));
        }
        for (int t = tid; t < D * D; t += DFU_NT) {
            const int i = t / D, jj = t % D;
            ival a = pt(0.0);
            for (int kk = 0; kk < N; kk++) dfu_ipfma(a, S.Ainv[i][kk], S.A[N + kk][jj]);
            S.K0[i][jj] = a;
            if (jj >= N) S.sAinv[i][jj - N] = dfu_spar(S.Ainv[i][jj]);
        }
        __syncthreads();
        int total = 1;
        for (int cq = 0; cq < N; cq++) total *= Ns;
        // ---- sub-boxes: G per warp at a time, the same trip count for every lane of a warp (shuffles inside) ----
        ival H[D];
        bool have = false;
        for (int base = warp * G; base < total; base += P) {
            const int pi = base + gl;
            const bool valid = act && pi < total;
            int rest = pi < total ? pi : total - 1;
            ival a = pt(0.0);
            {   // w_k = p_k + sum_m A[k][m] Up[m] - 1/2 over the sub-box: Up = U except the subdivided coordinates idx0 .. idx0 + n - 1
                ival Up[D];
#pragma unroll
                for (int i = 0; i < D; i++) Up[i] = S.U[i];
#pragma unroll
                for (int cq = 0; cq < N; cq++) { const ival pv = S.parts[cq][rest % Ns]; rest /= Ns;
#pragma unroll
                    for (int i = 0; i < D; i++) if (i == idx0 + cq) Up[i] = pv; }
#pragma unroll
                for (int m = 0; m < D; m++) dfu_ipfma(a, Up[m], S.A[k][m]);
            }
            const ival wk = isub(iadd(S.p[k], a), pt(0.5)), w2k = dfu_sqr(wk);
            const int kp = k + 1 < N ? k + 1 : N - 1, km = k > 0 ? k - 1 : 0;
            const ival wn = dfu_shfl(wk, lbase + kp), w2n = dfu_shfl(w2k, lbase + kp), w2p = dfu_shfl(w2k, lbase + km);
            // row k of M: diagonal Md_k and off-diagonal Mo_k = M[k][k+1]
            ival md = dfu_mul_sp(S.sb3[k], w2k);
            if (k > 0)     md = iadd(md, dfu_mul_sp(S.sc[km], w2p));
            if (k < N - 1) md = iadd(md, dfu_mul_sp(S.sc[k], w2n));
            md = isub(md, S.m0[k]);
            const ival mo = k < N - 1 ? dfu_mul_sp(S.sc2[k], imul(wk, wn)) : pt(0.0);
            ival Md[N], Mo[N];
#pragma unroll
            for (int kk = 0; kk < N; kk++) { Md[kk] = dfu_shfl(md, lbase + kk); Mo[kk] = dfu_shfl(mo, lbase + kk); }
            // column j of M A[:n,:] and of Ainv Df A
            ival MR[N];
#pragma unroll
            for (int kk = 0; kk < N; kk++) {
                ival m = pt(0.0);
                if (kk > 0)     dfu_ipfma(m, Mo[kk > 0 ? kk - 1 : 0], S.A[kk > 0 ? kk - 1 : 0][j]);
                dfu_ipfma(m, Md[kk], S.A[kk][j]);
                if (kk < N - 1) dfu_ipfma(m, Mo[kk], S.A[kk + 1 < N ? kk + 1 : kk][j]);
                MR[kk] = m;
            }
#pragma unroll
            for (int i = 0; i < D; i++) {
                ival h = S.K0[i][j];
#pragma unroll
                for (int kk = 0; kk < N; kk++) h = iadd(h, dfu_mul_sp(S.sAinv[i][kk], MR[kk]));
                if (valid) H[i] = have ? ihull(H[i], h) : h;
            }
            have = have || valid;
        }
        if (act) {
#pragma unroll
            for (int i = 0; i < D; i++) S.red[warp * G + gl][i][j] = have ? H[i] : mk(__longlong_as_double(0x7ff0000000000000LL), __longlong_as_double(0xfff0000000000000LL));   // empty: [+inf, -inf]
        }
        __syncthreads();
        for (int t = tid; t < D * D; t += DFU_NT) {
            const int i = t / D, jj = t % D;
            ival a = S.red[0][i][jj];
            for (int g2 = 1; g2 < P; g2++) a = ihull(a, S.red[g2][i][jj]);
            r.DF[i * CCP_MAX_DIM + jj].lo = a.lo; r.DF[i * CCP_MAX_DIM + jj].hi = a.hi;
        }
        if (tid == 0) { r.status = CCP_OK; r.parts = total; }
    }
}

} // namespace ccp
