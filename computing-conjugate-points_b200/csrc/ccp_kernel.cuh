// ccp_kernel.cuh -- the fused frame-propagation + conjugate-point-counting kernel (sm_100a, FP64).
//
// Replaces, for a whole batch of fronts in ONE launch, the reference's
//   propagateManifold::frameDet -> computeTotalTrajectory x n -> getTotalTrajectory -> topFrame
//   (propagateManifold.cpp:84-188, utils.cpp:175-263, topFrame.cpp:4-108,193-351).
// One warp owns one front.  All Taylor tables of the current step live in that warp's shared memory
// slice; set state is in shared memory too (small); per-lane accumulators in registers.  The sub-interval
// determinants, Jacobi derivative and the countZeros decision tree are evaluated by the lanes that
// produced the frame enclosures, so no trajectory (43.7 MB per front in the reference)
// is ever materialised.  Algorithm = oracle/c1.hpp (structured C^1 Lohner method), operation for operation.
#pragma once
#include "ccp_interval.cuh"
#include "../../include/ccp.h"

namespace ccp {

constexpr int KMAX = CCP_MAX_ORDER + 1;
constexpr int GMAX = CCP_MAX_GRID;

template <int N>
struct FrontSmem {
    static constexpr int D = 2 * N;
    static constexpr int NE = (N > 1) ? (N - 1) : 1;
    // Taylor tables: [set 0 = centre, 1 = hull]
    mt u[2][N][KMAX], v[2][N][KMAX], g[2][N][KMAX], q[2][NE][KMAX];
    mt w0[2][N], z0[2][N];
    mt Md[N][KMAX], Mo[NE][KMAX];
    mt Pu[N][D][KMAX], Pv[N][D][KMAX];
    ival Fk[KMAX][N][N];
    // set state (double buffered)
    double x[2][D], C[2][D][D];
    ival r[2][D], P[2][D][D];
    ival r0[D], Err[N][N], X[D], y[D], J[D][D], JC[D][D], Ws[D][N], bpar[N], cpar[NE];
    mt Wm[D][N];
    double rhoF[N], rhoFd[N];
    ival FD[GMAX][N][N];
    ccp_front_desc desc;
};

static_assert(sizeof(ccp_front_desc) % 16 == 0, "descriptors are staged with 16-byte loads");
static_assert(sizeof(ccp_front_result) % 16 == 0, "result records are 16-byte aligned");

// ---- determinants (oracle/common.hpp, same expression order) ----
template <int N> struct Det;
template <> struct Det<1> { __device__ static __forceinline__ ival det(const ival (&F)[1][1]) { return F[0][0]; } };
template <> struct Det<2> {
    __device__ static __forceinline__ ival det(const ival (&F)[2][2]) { return isub(imul(F[0][0], F[1][1]), imul(F[0][1], F[1][0])); }
};
__device__ __forceinline__ ival det2(const ival& a, const ival& b, const ival& c, const ival& d) { return isub(imul(a, d), imul(b, c)); }
template <> struct Det<3> {
    __device__ static __forceinline__ ival det(const ival (&m)[3][3]) {
        ival t0 = imul(m[0][0], det2(m[1][1], m[1][2], m[2][1], m[2][2]));
        ival t1 = imul(m[0][1], det2(m[1][0], m[1][2], m[2][0], m[2][2]));
        ival t2 = imul(m[0][2], det2(m[1][0], m[1][1], m[2][0], m[2][1]));
        return iadd(isub(t0, t1), t2);
    }
};
template <int N>
__device__ __forceinline__ ival minor_det(const ival (&F)[N][N], int ih, int jh) {
    ival Mn[N - 1][N - 1];
#pragma unroll
    for (int i = 0; i < N - 1; i++)
#pragma unroll
        for (int j = 0; j < N - 1; j++) Mn[i][j] = F[i < ih ? i : i + 1][j < jh ? j : j + 1];
    return Det<N - 1>::det(Mn);
}
template <> struct Det<4> {
    __device__ static __forceinline__ ival det(const ival (&F)[4][4]) {
        ival acc = imul(F[0][0], minor_det<4>(F, 0, 0));
#pragma unroll
        for (int j = 1; j < 4; j++) {
            ival term = imul(F[0][j], minor_det<4>(F, 0, j));
            acc = (j & 1) ? isub(acc, term) : iadd(acc, term);
        }
        return acc;
    }
};
template <int N>
__device__ __forceinline__ ival jacobi_derivative(const ival (&A)[N][N], const ival (&Ad)[N][N]) {
    ival tr = pt(0.0);
#pragma unroll
    for (int i = 0; i < N; i++) {
        ival diag = pt(0.0);
#pragma unroll
        for (int j = 0; j < N; j++) {
            ival cof = minor_det<N>(A, j, i);
            if ((i + j) & 1) cof = ineg(cof);
            diag = iadd(diag, imul(cof, Ad[j][i]));
        }
        tr = iadd(tr, diag);
    }
    return tr;
}

// decision of topFrame::countZeros for one sub-interval: 0 = certified no zero, 1 = certified zero, 2 = failure
__device__ __forceinline__ int count_decide(const ival& time, const ival& L, const ival& R, const ival& V, const ival& Dv) {
    ival LR = imul(L, R);
    if (LR.lo > 0) {
        if (idefinite(V)) return 0;
        if (idefinite(Dv)) return 0;
        ival w = isub(time, pt(time.lo));
        ival mvt = imul(w, Dv);
        ival half = mk(mvt.lo * 0.5, mvt.hi * 0.5);
        ival bound = ihull(iadd(L, half), isub(R, half));
        return idefinite(bound) ? 0 : 2;
    } else if (LR.hi < 0) {
        return idefinite(Dv) ? 1 : 2;
    }
    return 2;
}

struct Majorant { double rho_phi, rho_phi_d, rho_psi, rho_psi_d; ival Gs_u, Gs_v; bool ok; };

// oracle/c1.hpp:c1_majorant -- a-priori box + scalar majorant series (round-up), once per front, all lanes redundantly
template <int N>
__device__ inline Majorant majorant(const ival* b, const ival* c, int p, double h) {
    Majorant M;
    double bh = 0, cs = 0;
    for (int i = 0; i < N; i++) bh = dmax(bh, imag(b[i]));
    for (int i = 0; i < N; i++) {
        double s = 0;
        if (i > 0) s = __dadd_ru(s, imag(c[i - 1]));
        if (i < N - 1) s = __dadd_ru(s, imag(c[i]));
        cs = dmax(cs, s);
    }
    const double W0 = 0.5625, G0 = 0.25, V0 = 0.5;
    double bc = __dadd_ru(bh, cs);
    double F0 = __dmul_ru(__dmul_ru(bc, G0), W0);
    M.Gs_u = mk(__dadd_ru(-0.0625, __dmul_ru(h, V0)), __dadd_rd(1.0625, -__dmul_ru(h, V0)));
    M.Gs_v = mk(__dadd_ru(-0.5, __dmul_ru(h, F0)), __dadd_rd(0.5, -__dmul_ru(h, F0)));
    double w[CCP_MAX_ORDER + 3], g[CCP_MAX_ORDER + 3], be[CCP_MAX_ORDER + 3], Mh[CCP_MAX_ORDER + 3];
    w[0] = W0; g[0] = G0; w[1] = dmax(V0, F0);
    for (int k = 1; k <= p; k++) {
        double s = 0; for (int j = 0; j <= k; j++) s = __fma_ru(w[j], w[k - j], s);
        g[k] = s;
        double f = 0; for (int j = 0; j <= k; j++) f = __fma_ru(g[j], w[k - j], f);
        f = __dmul_ru(bc, f);
        w[k + 1] = __ddiv_ru(dmax(w[k], f), (double)(k + 1));
    }
    double hp = 1.0; for (int k = 0; k < p; k++) hp = __dmul_ru(hp, h);
    M.rho_phi   = __dmul_ru(w[p + 1], __dmul_ru(hp, h));
    M.rho_phi_d = __dmul_ru(__dmul_ru((double)(p + 1), w[p + 1]), hp);
    double M0 = __dadd_ru(__dadd_ru(__dmul_ru(bh, 0.75), __dmul_ru(cs, 0.25)), __dmul_ru(__dmul_ru(2.0, cs), __dmul_ru(W0, W0)));
    double a = dmax(1.0, M0), ah = __dmul_ru(a, h);
    be[0] = __dadd_ru(__dadd_ru(1.0, ah), __dmul_ru(ah, ah));
    Mh[0] = M0;
    double k3 = __dmul_ru(3.0, bc);
    for (int k = 1; k <= p; k++) Mh[k] = __dmul_ru(k3, g[k]);
    for (int k = 0; k <= p; k++) {
        double s = 0; for (int j = 0; j <= k; j++) s = __fma_ru(Mh[j], be[k - j], s);
        be[k + 1] = __ddiv_ru(dmax(be[k], s), (double)(k + 1));
    }
    M.rho_psi   = __dmul_ru(be[p + 1], __dmul_ru(hp, h));
    M.rho_psi_d = __dmul_ru(__dmul_ru((double)(p + 1), be[p + 1]), hp);
    M.ok = (ah <= 1.0) && isfinite(M.rho_phi) && isfinite(M.rho_psi) && M.Gs_u.lo < M.Gs_u.hi && M.Gs_v.lo < M.Gs_v.hi;
    return M;
}

__device__ __forceinline__ ival eval_pt(const mt* s, int p, double tau) {
    ival acc = mt2iv(s[p]);
    for (int k = p - 1; k >= 0; k--) acc = horner_pt(acc, tau, mt2iv(s[k]));
    return acc;
}
__device__ __forceinline__ ival eval_rg(const mt* s, int p, double tl, double th) {
    ival acc = mt2iv(s[p]);
    for (int k = p - 1; k >= 0; k--) acc = horner_rg(acc, tl, th, mt2iv(s[k]));
    return acc;
}
__device__ __forceinline__ ival eval_rg_deriv(const mt* s, int p, double tl, double th) {
    ival acc = imulk(mt2iv(s[p]), p);
    for (int k = p - 1; k >= 1; k--) acc = horner_rg(acc, tl, th, imulk(mt2iv(s[k]), k));
    return acc;
}
__device__ __forceinline__ ccp_interval toc(const ival& a) { ccp_interval o; o.lo = a.lo; o.hi = a.hi; return o; }
__device__ __forceinline__ ival fromc(const ccp_interval& a) { return mk(a.lo, a.hi); }

// host-side twin of oracle/common.hpp:step_count
__host__ __device__ inline int step_count(double L_minus, double h) {
    int S = (int)ceil(L_minus / h);
    while ((double)(S - 1) * h >= L_minus) S--;
    while ((double)S * h < L_minus) S++;
    return S;
}

// ------------------------------------------------------------------------------------------------
// One warp = one front.  `series` (optional) receives steps*grid records.
// ------------------------------------------------------------------------------------------------
template <int N>
__device__ void run_front(FrontSmem<N>& S, const ccp_front_desc* __restrict__ gdesc, ccp_front_result* __restrict__ gres,
                          ccp_series_rec* __restrict__ series, int series_cap) {
    constexpr int D = 2 * N;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;

    // ---- stage the descriptor: coalesced 16-byte loads HBM -> shared ----
    {
        const int4* src = reinterpret_cast<const int4*>(gdesc);
        int4* dst = reinterpret_cast<int4*>(&S.desc);
        for (int i = lane; i < (int)(sizeof(ccp_front_desc) / 16); i += 32) dst[i] = src[i];
    }
    __syncwarp();
    const ccp_front_desc& d = S.desc;
    const int p = d.order, g = d.grid;
    // result header defaults (lane 0 writes at the end)
    int status = CCP_OK;
    if (d.n != N || p < 2 || p > CCP_MAX_ORDER || g < 1 || g > GMAX || d.step_log2 < 1 || d.step_log2 > 20 || !(d.L_minus > 0) || !isfinite(d.L_minus))
        status = CCP_ST_BAD_DESC;
    if (status != CCP_OK) {
        if (lane == 0) {
            ccp_front_result z; memset(&z, 0, sizeof(z)); z.status = status; z.first_failure = -1;
            *gres = z;
        }
        return;
    }
    const double h = ldexp(1.0, -d.step_log2);
    if (lane < N) S.bpar[lane] = fromc(d.b[lane]);
    if (lane < N - 1) S.cpar[lane] = fromc(d.c[lane]);
    if (lane < D) {
        ival pi = fromc(d.p_i[lane]);
        double xm = mid_ru(pi);
        S.x[0][lane] = xm; S.r[0][lane] = isub(pi, pt(xm)); S.r0[lane] = fromc(d.Uxy[lane]);
    }
    for (int t = lane; t < D * D; t += 32) { int i = t / D, j = t % D; double a = d.A_u[i * CCP_MAX_DIM + j]; S.C[0][i][j] = a; S.P[0][i][j] = pt(a); }
    for (int t = lane; t < N * N; t += 32) { int j = t / N, k = t % N; S.Err[j][k] = fromc(d.Eu_err[j * CCP_MAX_N + k]); }
    // zero the parts of the result record that are filled piecemeal
    for (int t = lane; t < (int)(sizeof(ccp_front_result) / 8); t += 32) reinterpret_cast<double*>(gres)[t] = 0.0;
    __syncwarp();
    const Majorant mj = majorant<N>(S.bpar, S.cpar, p, h);
    if (!mj.ok) status = CCP_ST_ENCLOSURE;

    const int Stot = step_count(d.L_minus, h);
    int zero_count = 0, failure_count = 0, first_failure = -1, n_conj = 0;
    ival tprev = pt(0.0);
    int cur = 0;
    ival det_first = pt(0.0), det_last = pt(0.0);

    for (int s = 0; s < Stot && status == CCP_OK; s++) {
        const bool last = (s == Stot - 1);
        const double hs = last ? __dadd_ru(d.L_minus, -tprev.hi) : h;
        const int nxt = cur ^ 1;
        // ---- hull of the phi-set ----
        if (lane < D) {
            ival a = pt(S.x[cur][lane]);
            for (int j = 0; j < D; j++) a = iadd(a, imul(pt(S.C[cur][lane][j]), S.r0[j]));
            S.X[lane] = iadd(a, S.r[cur][lane]);
        }
        __syncwarp();
        {
            bool inside = true, fin = true;
            if (lane < D) { inside = isubset(S.X[lane], lane < N ? mj.Gs_u : mj.Gs_v); fin = ifinite(S.X[lane]); }
            if (!__all_sync(FULL, fin)) { status = CCP_ST_NONFINITE; break; }
            if (!__all_sync(FULL, inside)) { status = CCP_ST_ENCLOSURE; break; }
        }
        // ---- Taylor tables of phi: set 0 = centre x, set 1 = hull X ----
        if (lane < 2 * N) {
            const int set = lane / N, i = lane % N;
            ival u0 = set ? S.X[i] : pt(S.x[cur][i]);
            ival v0 = set ? S.X[N + i] : pt(S.x[cur][N + i]);
            S.u[set][i][0] = iv2mt(u0); S.v[set][i][0] = iv2mt(v0);
            S.w0[set][i] = iv2mt(isub(u0, pt(0.5))); S.z0[set][i] = iv2mt(isub(pt(1.0), u0));
        }
        // Psi order 0 = identity
        for (int t = lane; t < N * D; t += 32) {
            const int i = t / D, cc = t % D;
            S.Pu[i][cc][0] = iv2mt(pt(cc == i ? 1.0 : 0.0));
            S.Pv[i][cc][0] = iv2mt(pt(cc == N + i ? 1.0 : 0.0));
        }
        __syncwarp();
        for (int k = 0; k < p; k++) {
            // stage A: g_i[k] = (u*z)[k], q_e[k] = (w_e*w_{e+1})[k]
            if (lane < 2 * N) {
                const int set = lane / N, i = lane % N;
                const mt* uu = S.u[set][i];
                acc4 a = acc_zero();
                for (int j = 0; j < k; j++) acc_term_neg(a, uu[j], uu[k - j]);
                acc_term(a, uu[k], S.z0[set][i]);
                S.g[set][i][k] = iv2mt(acc_finish(a));
            } else if (lane < 2 * N + 2 * (N - 1)) {
                const int t = lane - 2 * N, set = t / (N - 1), e = t % (N - 1);
                const mt* ua = S.u[set][e]; const mt* ub = S.u[set][e + 1];
                acc4 a = acc_zero();
                for (int j = 0; j <= k; j++) acc_term(a, j == 0 ? S.w0[set][e] : ua[j], k - j == 0 ? S.w0[set][e + 1] : ub[k - j]);
                S.q[set][e][k] = iv2mt(acc_finish(a));
            }
            __syncwarp();
            // stage B: v_i[k+1], u_i[k+1]  (lanes 0..2N-1)  and  Md_i[k], Mo_e[k] from the hull (lanes 2N..)
            if (lane < 2 * N) {
                const int set = lane / N, i = lane % N;
                const mt* uu = S.u[set][i]; const mt w0 = S.w0[set][i];
                acc4 a = acc_zero();
                { const mt* gg = S.g[set][i]; for (int j = 0; j <= k; j++) acc_term(a, gg[j], k - j == 0 ? w0 : uu[k - j]); }
                ival f = ineg(imul(S.bpar[i], acc_finish(a)));
                if (i > 0)     { acc4 a2 = acc_zero(); const mt* gg = S.g[set][i - 1]; for (int j = 0; j <= k; j++) acc_term(a2, gg[j], k - j == 0 ? w0 : uu[k - j]); f = isub(f, imul(S.cpar[i - 1], acc_finish(a2))); }
                if (i < N - 1) { acc4 a2 = acc_zero(); const mt* gg = S.g[set][i + 1]; for (int j = 0; j <= k; j++) acc_term(a2, gg[j], k - j == 0 ? w0 : uu[k - j]); f = isub(f, imul(S.cpar[i], acc_finish(a2))); }
                S.v[set][i][k + 1] = iv2mt(idivk(f, k + 1));
                S.u[set][i][k + 1] = iv2mt(idivk(mt2iv(S.v[set][i][k]), k + 1));
            } else if (lane < 3 * N) {
                const int i = lane - 2 * N;
                ival t3 = imulk(mt2iv(S.g[1][i][k]), 3);
                if (k == 0) t3 = isub(t3, pt(0.5));
                ival dd = ineg(imul(S.bpar[i], t3));
                if (i > 0)     dd = isub(dd, imul(S.cpar[i - 1], mt2iv(S.g[1][i - 1][k])));
                if (i < N - 1) dd = isub(dd, imul(S.cpar[i], mt2iv(S.g[1][i + 1][k])));
                S.Md[i][k] = iv2mt(dd);
            } else if (lane < 3 * N + (N - 1)) {
                const int e = lane - 3 * N;
                ival c2 = mk(2.0 * S.cpar[e].lo, 2.0 * S.cpar[e].hi);
                S.Mo[e][k] = iv2mt(imul(c2, mt2iv(S.q[1][e][k])));
            }
            __syncwarp();
            // stage C: Psi order k+1
            for (int t = lane; t < N * D; t += 32) {
                const int i = t / D, cc = t % D;
                acc4 a = acc_zero();
                { const mt* mm = S.Md[i]; const mt* pp = S.Pu[i][cc]; for (int j = 0; j <= k; j++) acc_term(a, mm[j], pp[k - j]); }
                if (i > 0)     { const mt* mm = S.Mo[i - 1]; const mt* pp = S.Pu[i - 1][cc]; for (int j = 0; j <= k; j++) acc_term(a, mm[j], pp[k - j]); }
                if (i < N - 1) { const mt* mm = S.Mo[i]; const mt* pp = S.Pu[i + 1][cc]; for (int j = 0; j <= k; j++) acc_term(a, mm[j], pp[k - j]); }
                S.Pv[i][cc][k + 1] = iv2mt(idivk(acc_finish(a), k + 1));
                S.Pu[i][cc][k + 1] = iv2mt(idivk(mt2iv(S.Pv[i][cc][k]), k + 1));
            }
            __syncwarp();
        }
        // ---- frame at step start, frame tables F_k = Psi^u_k * Ws ----
        for (int t = lane; t < D * N; t += 32) {
            const int cc = t / N, bb = t % N;
            ival a = S.P[cur][cc][bb];
            for (int j = 0; j < N; j++) a = iadd(a, imul(S.P[cur][cc][N + j], S.Err[j][bb]));
            S.Ws[cc][bb] = a; S.Wm[cc][bb] = iv2mt(a);
        }
        __syncwarp();
        if (lane < N) {
            double sm = 0; for (int cc = 0; cc < D; cc++) sm = __dadd_ru(sm, imag(S.Ws[cc][lane]));
            S.rhoF[lane] = __dmul_ru(mj.rho_psi, sm); S.rhoFd[lane] = __dmul_ru(mj.rho_psi_d, sm);
        }
        __syncwarp();
        for (int t = lane; t < (p + 1) * N * N; t += 32) {
            const int k = t / (N * N), a = (t / N) % N, bb = t % N;
            acc4 ac = acc_zero();
            for (int cc = 0; cc < D; cc++) acc_term(ac, S.Pu[a][cc][k], S.Wm[cc][bb]);
            ival f = acc_finish(ac);
            if (k == 0) f = iadd(f, sym(S.rhoF[bb]));
            S.Fk[k][a][bb] = f;
        }
        __syncwarp();
        // ---- grid loop + topFrame: lanes 0..g evaluate the frame at the grid points (and its derivative range for
        //      sub-interval `lane`), lanes 16..16+g-1 the range enclosure, dets and decisions ----
        {
            const int gi = lane & 15;                                  // grid index handled by this lane
            const bool is_point = lane <= g && lane < 16;              // lanes 0..g
            const bool is_range = lane >= 16 && (lane - 16) < g;       // lanes 16..16+g-1  -> sub-interval gi
            double tau_lo = 0, tau_hi = 0;
            if (is_point) { tau_lo = (gi < g) ? __ddiv_rd(__dmul_rd((double)gi, hs), (double)g) : hs; tau_hi = (gi + 1 < g) ? __ddiv_rd(__dmul_rd((double)(gi + 1), hs), (double)g) : hs; }
            if (is_range) { tau_lo = __ddiv_rd(__dmul_rd((double)gi, hs), (double)g); tau_hi = (gi + 1 < g) ? __ddiv_rd(__dmul_rd((double)(gi + 1), hs), (double)g) : hs; }
            ival Fm[N][N];
            ival detP = pt(0.0);
            if (is_point) {
#pragma unroll
                for (int a = 0; a < N; a++)
#pragma unroll
                    for (int bb = 0; bb < N; bb++) {
                        ival acc = S.Fk[p][a][bb];
                        for (int k = p - 1; k >= 0; k--) acc = horner_pt(acc, tau_lo, S.Fk[k][a][bb]);
                        Fm[a][bb] = acc;
                    }
                detP = Det<N>::det(Fm);
                if (gi < g) {          // derivative range on sub-interval gi, handed to the range lane through shared memory
#pragma unroll
                    for (int a = 0; a < N; a++)
#pragma unroll
                        for (int bb = 0; bb < N; bb++) {
                            ival acc = imulk(S.Fk[p][a][bb], p);
                            for (int k = p - 1; k >= 1; k--) acc = horner_rg(acc, tau_lo, tau_hi, imulk(S.Fk[k][a][bb], k));
                            S.FD[gi][a][bb] = iadd(acc, sym(S.rhoFd[bb]));
                        }
                }
            } else if (is_range) {
#pragma unroll
                for (int a = 0; a < N; a++)
#pragma unroll
                    for (int bb = 0; bb < N; bb++) {
                        ival acc = S.Fk[p][a][bb];
                        for (int k = p - 1; k >= 0; k--) acc = horner_rg(acc, tau_lo, tau_hi, S.Fk[k][a][bb]);
                        Fm[a][bb] = acc;
                    }
            }
            __syncwarp();
            // det at left / right end of sub-interval gi comes from point lanes gi and gi+1
            const double dLlo = __shfl_sync(FULL, detP.lo, gi), dLhi = __shfl_sync(FULL, detP.hi, gi);
            const double dRlo = __shfl_sync(FULL, detP.lo, (gi + 1) & 15), dRhi = __shfl_sync(FULL, detP.hi, (gi + 1) & 15);
            int decision = 0;
            ival timeI = pt(0.0), dV = pt(0.0), dD = pt(0.0);
            if (is_range) {
                ival FDm[N][N];
#pragma unroll
                for (int a = 0; a < N; a++)
#pragma unroll
                    for (int bb = 0; bb < N; bb++) FDm[a][bb] = S.FD[gi][a][bb];
                dV = Det<N>::det(Fm);
                dD = jacobi_derivative<N>(Fm, FDm);
                timeI = mk(__dadd_rd(tprev.lo, tau_lo), __dadd_ru(tprev.hi, tau_hi));
                const ival dL = mk(dLlo, dLhi), dR = mk(dRlo, dRhi);
                decision = count_decide(timeI, dL, dR, dV, dD);
                if (!(ifinite(dL) && ifinite(dR) && ifinite(dV) && ifinite(dD))) decision |= 4;
                const int idx = s * g + gi;
                if (series != nullptr && idx < series_cap) {
                    ccp_series_rec rec; rec.time = toc(timeI); rec.det_L = toc(dL); rec.det_R = toc(dR); rec.det_V = toc(dV); rec.det_D = toc(dD);
                    series[idx] = rec;
                }
                if (idx == 0) det_first = dV;
                if (idx == Stot * g - 1) det_last = dV;
            }
            const unsigned zmask = __ballot_sync(FULL, is_range && (decision & 3) == 1);
            const unsigned fmask = __ballot_sync(FULL, is_range && (decision & 3) == 2);
            const unsigned nmask = __ballot_sync(FULL, is_range && (decision & 4));
            if (nmask) status = CCP_ST_NONFINITE;
            if (fmask && failure_count == 0) first_failure = s * g + (__ffs(fmask) - 1 - 16);
            failure_count += __popc(fmask);
            if (zmask) {
                if (is_range && (decision & 3) == 1) {
                    const int slot = n_conj + __popc(zmask & ((1u << lane) - 1u));
                    if (slot < CCP_MAX_CONJ) { gres->conj_time[slot] = toc(timeI); gres->conj_index[slot] = s * g + gi; }
                }
                n_conj += __popc(zmask); zero_count += __popc(zmask);
            }
            // first / last det range enclosure live on one lane: broadcast
            det_first.lo = __shfl_sync(FULL, det_first.lo, 16); det_first.hi = __shfl_sync(FULL, det_first.hi, 16);
            if (last) { det_last.lo = __shfl_sync(FULL, det_last.lo, 16 + g - 1); det_last.hi = __shfl_sync(FULL, det_last.hi, 16 + g - 1); }
        }
        // ---- first-frame data (topFrame::getFirstFrame) ----
        if (s == 0) {
            const double t0 = 0.0, t1 = (1 < g) ? __ddiv_rd(__dmul_rd(1.0, hs), (double)g) : hs;
            for (int t = lane; t < D * N; t += 32) { const int rr = t / N, k = t % N; gres->first_frame[rr * (CCP_MAX_N + 1) + k] = toc(S.Ws[rr][k]); }
            if (lane < N) {
                ival du = iadd(eval_rg(S.v[1][lane], p, t0, t1), sym(mj.rho_phi));
                ival dv = iadd(eval_rg_deriv(S.v[1][lane], p, t0, t1), sym(mj.rho_phi_d));
                gres->first_frame[lane * (CCP_MAX_N + 1) + N] = toc(du);
                gres->first_frame[(N + lane) * (CCP_MAX_N + 1) + N] = toc(dv);
            }
        }
        // ---- step: y = Phi(x) (centre tables), J = DPhi over the hull (Psi tables) ----
        for (int t = lane; t < D + D * D; t += 32) {
            if (t < D) {
                const int i = t % N;
                S.y[t] = iadd(eval_pt(t < N ? S.u[0][i] : S.v[0][i], p, hs), sym(mj.rho_phi));
            } else {
                const int e = t - D, row = e / D, cc = e % D, i = row % N;
                S.J[row][cc] = iadd(eval_pt(row < N ? S.Pu[i][cc] : S.Pv[i][cc], p, hs), sym(mj.rho_psi));
            }
        }
        __syncwarp();
        if (last) {   // topFrame::getLastFrame
            constexpr int ld = CCP_MAX_N + 2;
            for (int t = lane; t < D * N; t += 32) {
                const int rr = t / N, k = t % N;
                ival a = pt(0.0);
                for (int cc = 0; cc < D; cc++) a = iadd(a, imul(S.J[rr][cc], S.Ws[cc][k]));
                gres->last_frame[rr * ld + k] = toc(a);
            }
            if (lane < N) {
                const double tl = __ddiv_rd(__dmul_rd((double)(g - 1), hs), (double)g);
                ival du = iadd(eval_rg(S.v[1][lane], p, tl, hs), sym(mj.rho_phi));
                ival dv = iadd(eval_rg_deriv(S.v[1][lane], p, tl, hs), sym(mj.rho_phi_d));
                gres->last_frame[lane * ld + N] = toc(du);
                gres->last_frame[(N + lane) * ld + N] = toc(dv);
            }
        }
        // ---- set update ----
        if (lane < D) S.x[nxt][lane] = mid_ru(S.y[lane]);
        for (int t = lane; t < D * D; t += 32) {
            const int i = t / D, j = t % D;
            ival a = pt(0.0);
            for (int k = 0; k < D; k++) a = iadd(a, imul(S.J[i][k], pt(S.C[cur][k][j])));
            S.JC[i][j] = a; S.C[nxt][i][j] = mid_ru(a);
        }
        for (int t = lane; t < D * D; t += 32) {
            const int i = t / D, j = t % D;
            ival a = pt(0.0);
            for (int k = 0; k < D; k++) a = iadd(a, imul(S.J[i][k], S.P[cur][k][j]));
            S.P[nxt][i][j] = a;
        }
        __syncwarp();
        if (lane < D) {
            const int i = lane;
            ival z = isub(S.y[i], pt(S.x[nxt][i]));
            for (int j = 0; j < D; j++) z = iadd(z, imul(isub(S.JC[i][j], pt(S.C[nxt][i][j])), S.r0[j]));
            ival a = pt(0.0);
            for (int j = 0; j < D; j++) a = iadd(a, imul(S.J[i][j], S.r[cur][j]));
            S.r[nxt][i] = iadd(a, z);
        }
        __syncwarp();
        tprev = mk(__dadd_rd(tprev.lo, hs), __dadd_ru(tprev.hi, hs));
        cur = nxt;
        if (last && lane < D) {
            constexpr int ld = CCP_MAX_N + 2;
            ival a = pt(S.x[cur][lane]);
            for (int j = 0; j < D; j++) a = iadd(a, imul(pt(S.C[cur][lane][j]), S.r0[j]));
            gres->last_frame[lane * ld + N + 1] = toc(iadd(a, S.r[cur][lane]));
        }
    }
    __syncwarp();
    if (lane == 0) {
        gres->status = status; gres->zero_count = zero_count; gres->failure_count = failure_count;
        gres->series_length = Stot * g; gres->steps = Stot; gres->n_conj = n_conj < CCP_MAX_CONJ ? n_conj : CCP_MAX_CONJ;
        gres->first_failure = first_failure; gres->reserved = 0;
        gres->det_first = toc(det_first); gres->det_last = toc(det_last);
    }
}

template <int N>
__global__ void __launch_bounds__(32) front_kernel(const ccp_front_desc* __restrict__ descs, ccp_front_result* __restrict__ results,
                                                   const int* __restrict__ index, int count, ccp_series_rec* series, int series_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrontSmem<N>& S = *reinterpret_cast<FrontSmem<N>*>(smem_raw);
    for (int w = blockIdx.x; w < count; w += gridDim.x) {
        const int f = index ? index[w] : w;
        run_front<N>(S, descs + f, results + f, series, series_cap);
        __syncwarp();
    }
}

} // namespace ccp
