// ccp_kernel.cuh -- the fused frame-propagation + conjugate-point-counting kernel (sm_100a, FP64).
//
// Replaces, for a whole batch of fronts in ONE launch, the reference's
//   propagateManifold::frameDet -> computeTotalTrajectory x n -> getTotalTrajectory -> topFrame
//   (propagateManifold.cpp:84-188, utils.cpp:175-263, topFrame.cpp:4-108,193-351).
// One CTA of 64 threads (two specialised warps) owns one front.  All Taylor tables of the current step live in that CTA's shared
// memory; set state is in shared memory too (small); per-lane accumulators in registers.  The sub-interval determinants, Jacobi
// derivative and the countZeros decision tree are evaluated in the same kernel, so no trajectory (43.7 MB per front in the
// reference) is ever materialised.  Algorithm = oracle/c2.hpp (structured C^1 Lohner method with first-order tracking
// of the frame's dependence on the front's box and centres advanced by increments), operation for operation.
#pragma once
#include "ccp_interval.cuh"
#include <cuda/std/type_traits>
#include "../../include/ccp.h"

#ifndef CCP_STAGE0
#define CCP_STAGE0 2            // 2: the first stage (J = 0) has its own compact instantiation; 0: it runs through the general one
#endif
#ifndef CCP_SINGLES
#define CCP_SINGLES 1           // 1: one variational product per lane, two columns of Psi hosted by the idle lanes of the phi warp (n <= 3)
#endif
#ifndef CCP_UNROLL_S
#define CCP_UNROLL_S 1      // unrolling of the two-term loops of the Cauchy passes (single tasks / pairs of tasks)
#endif
#ifndef CCP_UNROLL_P
#define CCP_UNROLL_P 1      // 1 measured best at full occupancy: the step loop streams ~78 KB of code per step (L1.5 instruction cache: 32 KB), smaller is faster
#endif
#ifndef CCP_UNROLL_H
#define CCP_UNROLL_H 4      // Horner chains of the step image
#endif
#ifndef CCP_UNROLL_G
#define CCP_UNROLL_G 4      // Horner chains of the grid loop (4: -1 % against 2 with the v26 code, 1 and 8 are slower; profiles/r2_v26_unroll_experiments.txt)
#endif

namespace ccp {
constexpr int UNROLL_S = CCP_UNROLL_S, UNROLL_P = CCP_UNROLL_P, UNROLL_H = CCP_UNROLL_H, UNROLL_G = CCP_UNROLL_G;

// Optional per-phase cycle counters (make prof): block 0 adds clock64() deltas into g_prof[phase].
#ifdef CCP_TRACE
// timeline trace of ONE step of block 0: (clock, marker) pairs written by lane 0 of each warp (scripts/gpu_trace.py)
__device__ long long g_trace[2][512];
__device__ int g_trace_n[2];
#define TRC_DECL() int trc_n = 0
#define TRC(id) do { if (blockIdx.x == 0 && s == 500 && (threadIdx.x & 31) == 0 && trc_n < 255) { g_trace[threadIdx.x >> 5][2 * trc_n] = clock64(); g_trace[threadIdx.x >> 5][2 * trc_n + 1] = (id); trc_n++; g_trace_n[threadIdx.x >> 5] = trc_n; } } while (0)
#else
#define TRC_DECL() do {} while (0)
#define TRC(id) do {} while (0)
#endif
#ifdef CCP_PROFILE
__device__ unsigned long long g_prof[16];
#define PROF_T0() long long prof_t = clock64()
#define PROF(i) do { long long prof_n = clock64(); if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_prof[i], (unsigned long long)(prof_n - prof_t)); prof_t = prof_n; } while (0)
#else
#define PROF_T0() do {} while (0)
#define PROF(i) do {} while (0)
#endif

// ---- soft per-SM alignment of co-resident fronts -------------------------------------------------------------------------------
// The step loop is ~100 KB of straight-line SASS, far beyond the instruction caches (L0 6 KB, L1.5 32 KB): seven co-resident fronts
// that drift apart stream that code from L2 seven times per step (ncu: stalled_no_instruction 0.2 per issue with one front per SM,
// 1.3 with seven).  Fronts of one SM therefore meet once per step at a barrier in global memory, keyed by %smid, so that they walk
// through the code together and share the fetched lines.  The barrier is only a performance hint: membership is dynamic (a CTA
// registers when it starts and leaves when it is done), every wait is bounded by SM_ALIGN_TIMEOUT cycles, and any inconsistency
// (time-outs, kernels of two streams sharing an SM) heals at the next release.  Results never depend on it.
constexpr long long SM_ALIGN_TIMEOUT = 40000;
constexpr int SM_ALIGN_SLOTS = 512;
struct __align__(16) SmAlign { unsigned members, count, gen, pad; };       // one 16-byte record per SM: one L2 sector serves a whole wait
__device__ SmAlign g_sm_align[SM_ALIGN_SLOTS];
__device__ __forceinline__ unsigned sm_id() { unsigned v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v % SM_ALIGN_SLOTS; }
__device__ __forceinline__ void sm_align_release(SmAlign& a) { *(volatile unsigned*)&a.count = 0u; atomicAdd(&a.gen, 1u); }
__device__ __forceinline__ unsigned sm_align_join(unsigned sm) { return atomicAdd(&g_sm_align[sm].members, 1u); }
__device__ __forceinline__ void sm_align_leave(unsigned sm) {
    SmAlign& a = g_sm_align[sm];
    const unsigned left = atomicSub(&a.members, 1u) - 1u;
    if (left > 0u && *(volatile unsigned*)&a.count >= left) sm_align_release(a);      // the others may be waiting for this CTA
}
// called by ONE thread of the CTA; the CTA synchronises afterwards.  count = arrivals of the current generation; whoever completes it
// resets the count and opens the next generation.  A waiter that gives up simply goes on (its arrival stays counted, so the others are
// still released when the last one arrives).
__device__ __forceinline__ void sm_align_wait(unsigned sm) {
    SmAlign& a = g_sm_align[sm];
    const unsigned gen = *(volatile unsigned*)&a.gen;
    const unsigned members = *(volatile unsigned*)&a.members;
    if (members <= 1u) return;
    const unsigned arrived = atomicAdd(&a.count, 1u) + 1u;
    if (arrived >= members) { sm_align_release(a); return; }
    const long long t0 = clock64();
    while (*(volatile unsigned*)&a.gen == gen && clock64() - t0 < SM_ALIGN_TIMEOUT) { }
}

constexpr int KMAX = CCP_MAX_ORDER + 1;
constexpr int GMAX = CCP_MAX_GRID;
// [rd(1/k), ru(1/k)] and ru(ru(1/k) (1 + 2^-50)) (radius factor of the (mid,t) scaling by 1/k, mt_scale), k <= KMAX + 1: the same for
// every front, filled once per device by inv_table_kernel (ccp_engine.cu) with the device's own directed divisions
__constant__ double c_invlo[KMAX + 2], c_invhi[KMAX + 2], c_invs[KMAX + 2];
__global__ void inv_table_kernel(double* out) {
    const int k = threadIdx.x;
    if (k >= KMAX + 2) return;
    out[k] = k >= 1 ? __ddiv_rd(1.0, (double)k) : 0.0;
    out[KMAX + 2 + k] = k >= 1 ? __ddiv_ru(1.0, (double)k) : 0.0;
    out[2 * (KMAX + 2) + k] = k >= 1 ? __dmul_ru(__ddiv_ru(1.0, (double)k), 1.0000000000000009) : 0.0;
}

// Series index map of the table array T (one row = one Taylor series, KMAX coefficients, 16 B each).
// Row stride 21*16 B = 84 words => 8 consecutive rows hit 8 disjoint 4-bank groups: LDS.128 conflict-free.
template <int N> struct Lay {
    static constexpr int D = 2 * N, NE = N - 1;
    // ONE table set: the centre trajectory x (the dependence on the set is carried by the second variation, see run_front)
    __host__ __device__ static constexpr int U(int i) { return i; }
    __host__ __device__ static constexpr int V(int i) { return N + i; }
    __host__ __device__ static constexpr int G(int i) { return 2 * N + i; }
    __host__ __device__ static constexpr int Q(int e) { return 3 * N + e; }
    __host__ __device__ static constexpr int MD(int i) { return 3 * N + NE + i; }
    __host__ __device__ static constexpr int MO(int e) { return 4 * N + NE + e; }
    __host__ __device__ static constexpr int PU(int i, int c) { return 4 * N + 2 * NE + i * D + c; }
    __host__ __device__ static constexpr int PV(int i, int c) { return 4 * N + 2 * NE + N * D + i * D + c; }
    __host__ __device__ static constexpr int Z(int i) { return 4 * N + 2 * NE + 2 * N * D + i; }   // z = 1 - u
    static constexpr int NS = 5 * N + 2 * NE + 2 * N * D;
    // w = u - 1/2 differs from u only at order 0: w0 lives in the unused slot G(i)[KMAX-1] (g is only needed
    // to order p-1) and enters the products as a peeled end term.  z = 1 - u has its own rows (z_k = -u_k).
    static constexpr int NA = N + NE;                 // A tasks: g[i] = u*z, q[e] = w_e*w_{e+1}
    static constexpr int NB = N;                      // B tasks: H_i*w_i with H_i = -(b_i g_i + c_{i-1} g_{i-1} + c_i g_{i+1}) (rows G hold H): v' = F(u) = H w
    static constexpr int NC = D * (3 * N - 2);        // C tasks: Md_i*Pu_i^c, Mo*Pu_{i+-1}^c
    static constexpr int NTASK = NA + NB + NC;
    static constexpr int NPH = (NTASK + 31) / 32 + ((NA + NC) <= 32 && NB > 0 && NTASK <= 32 ? 1 : 0);   // >= 2 always (A before B)
};

struct Majorant { double rho_phi, rho_phi_d, rho_psi, rho_psi_d, g1, g5; ival Gs_u, Gs_v; bool ok; };   // g1, g5: second variation (majorant2)
// parameter interval as sign * [pl, ph] with 0 <= pl <= ph (see imul_par)
struct SPar { double pl, ph; int neg, straddle; };

// low-order Taylor data over the hull [X] of the phi-set (oracle/c2.hpp:HullM): W[q] = coefficients of u - 1/2 (q <= 3),
// G = u(1-u), Pw = w_0 w_1, dense tridiagonal M_0, M_1 -- all in (mid,t) form
template <int N> struct HullM { mt W[4][N], G[N], Pw[N], M0[N][N], M1[N][N]; };
// parameter tables of a front (oracle/c2.hpp:ParM); missing chain neighbours are zero parameters
template <int N> struct ParM { mt B[N], B3[N], B6[N], Cm[N], Cp[N], C2m[N], C2p[N], C2[N]; ival halfB[N]; mt PN[N][N][3]; };

// Overlay of the Taylor-table array once the tables of a step are dead (after the frame tables F_k are built).
template <int N> struct PostTab {
    static constexpr int D = 2 * N;
    ival Xi[D + 1][D][D];                    // second variation DPhi'[c_j] over the hull, j < D; Xi[D]: direction = the box r
    mt dJm[D][D], Jm[D][D], Jcm[D][D];       // (mid,t) forms of dJ, J, Jc for the set update
    double magXu[N][D], magXv[N][D];
    ival DT[3 * (GMAX + 1)];                 // determinants of the grid loop
    union {
        ival EV[(2 * GMAX + 1) * N * N];     // frame enclosures of the grid loop
        struct { ival pad[(2 * GMAX + 1) * N * N]; ival E[D][D][N]; double Thn[D][D][N], Tcn[D][N]; ival Rn[D][N], Psn[D][N]; } e;   // set update (behind EV, which the deciding warp may still read): (enclosure of Theta_j+ minus its midpoint) * r0_j
        struct {                             // transients of the second-variation phase (AM reuses WD, NN reuses DD)
            union { mt WD[D + 1][4][N][N]; double AM[D + 1][D][D]; };
            union { mt DD[D + 1][4][N]; mt NN[D + 1][4][N][N]; };
        } x;
    } u;
};

template <int N>
struct FrontSmem {
    static constexpr int D = 2 * N;
    static constexpr int NE = (N > 1) ? (N - 1) : 1;
    // Table array: Taylor tables of the centre trajectory.  Once they are dead, PostTab overlays rows [0, PTROWS) and the frame tables
    // Fk[KMAX][N][N] (exactly N*N rows) live in the LAST N*N rows -- rows the frame-table phase no longer reads (they hold Psi^v and z,
    // or lie beyond the tables), so building Fk there needs no extra copy.  This keeps a front of n = 3 below 28,160 B: 8 fronts per SM.
    static constexpr int PTROWS = (int)((sizeof(PostTab<N>) + KMAX * sizeof(mt) - 1) / (KMAX * sizeof(mt)));
    static constexpr int TROWS = PTROWS + N * N > Lay<N>::NS ? PTROWS + N * N : Lay<N>::NS;
    static constexpr int FKROW = TROWS - N * N;
    mt T[TROWS][KMAX];
    __device__ __forceinline__ ival (*fk())[N][N] { return reinterpret_cast<ival (*)[N][N]>(&T[FKROW][0]); }
    // front set x + C r0 + r (double buffered)
    double x[2][D], C[2][D][D];
    ival r[2][D];
    // unstable columns of Dphi_t A_u:  Tc + sum_j Th[j] r0_j + R  (Tc, Th points);  stable columns Ps (interval)
    double Tc[D][N], Th[D][D][N];
    ival R[D][N], Ps[D][N], S0[D][N];
    mt Rm[D][N], Tdm[D][N], Psm[D][N];       // (mid,t) forms of R, S0 + R, Ps (step start) for the set update
    ival r0[D], Err[N][N], X[D], y[D], Jc[D][D], Kc[D][D], J[D][D], Ws[D][N], bpar[N], cpar[NE];   // y, Kc: increments Phi(x) - x, DPhi(x) - I
    // During the table loop Jc, Kc, J are dead: they hold the second scratch set (order J+1 of a stage), a dummy slot and, from INV_OFF, the reciprocal table
    static constexpr int INV_OFF = 3 * D * D - (KMAX + 2);
    mt Wm[D][N];
    HullM<N> hm;
    ParM<N> pm;
    mt r0m[D + 1], Errm[N][N];               // (mid,t) forms; r0m[D] = 1 weights the direction "box r"
    double boxFd[N][N];
    double gp[GMAX + 1];                     // grid points of the current step (utils.cpp:219-223)
    int flag;
    int vswap;                               // 0 or 32: the two warps of the front exchange their roles (tid = threadIdx.x ^ vswap), see front_kernel
    int renorms;                             // basis renormalisations applied so far (renormalise)
    ival G[N][N], sc;                        // true frame = normalised frame * G ;  sc encloses det G > 0
    Majorant mj;
    double g15[2];
    struct { int zero_count, failure_count, first_failure, n_conj; ival tprev, det_first, det_last; } st;
    struct { int p, g, step_log2, Stot, renorm_mask, series_cap; double L_minus, h; const ccp_front_desc* gdesc; ccp_front_result* gres; ccp_series_rec* series; } cfg;      // the front's configuration: read where it is used instead of living in registers     // tallies of the front (warp 0) and the time reached: in shared memory, not in registers that would be live across the whole step
    union alignas(16) {
        ival scratch[Lay<N>::NB + Lay<N>::NC + 2 * N];   // table loop: Cauchy sums awaiting combination; then [0,0]; then the newest g, q
        mt JCr[2 * N][2 * N];                // set update: J*C - C+
        struct { double Hm[N][N]; ival Hi[N][N]; int ok; } rn;   // renormalise: H and the enclosure of its inverse
    } un;
};

static_assert(sizeof(PostTab<2>) <= sizeof(FrontSmem<2>::T) && sizeof(PostTab<3>) <= sizeof(FrontSmem<3>::T) && sizeof(PostTab<4>) <= sizeof(FrontSmem<4>::T),
              "PostTab must fit into the table array");
static_assert(FrontSmem<3>::TROWS == Lay<3>::NS, "n = 3: the overlay must not enlarge the table array");
static_assert(sizeof(FrontSmem<3>) <= 28160, "n = 3: 8 fronts per SM (228 KB of shared memory, 1 KB reserved per CTA)");
static_assert(FrontSmem<2>::FKROW > Lay<2>::PU(1, 3) && FrontSmem<3>::FKROW > Lay<3>::PU(2, 5) && FrontSmem<4>::FKROW > Lay<4>::PU(3, 7), "Fk must not overlay the Psi^u rows it is built from");
static_assert(offsetof(FrontSmem<3>, Kc) == offsetof(FrontSmem<3>, Jc) + sizeof(FrontSmem<3>::Jc) && offsetof(FrontSmem<3>, J) == offsetof(FrontSmem<3>, Kc) + sizeof(FrontSmem<3>::Kc), "Jc, Kc, J are contiguous");
static_assert(sizeof(ccp_front_desc) <= sizeof(FrontSmem<2>::T), "the descriptor is staged in the table array");
static_assert(sizeof(ccp_front_desc) % 16 == 0, "descriptors are staged with 16-byte loads");
static_assert(sizeof(ccp_front_result) % 16 == 0, "result records are 16-byte aligned");

// ---- determinants (oracle/common.hpp, same expression order) ----
template <int N> struct Det;
template <> struct Det<1> { __device__ static __forceinline__ ival det(const ival (&F)[1][1]) { return F[0][0]; } };
template <> struct Det<2> {
    __device__ static __forceinline__ ival det(const ival (&F)[2][2]) { return isub(imul_nl(F[0][0], F[1][1]), imul_nl(F[0][1], F[1][0])); }
};
__device__ __forceinline__ ival det2(const ival& a, const ival& b, const ival& c, const ival& d) { return isub(imul_nl(a, d), imul_nl(b, c)); }
template <> struct Det<3> {
    __device__ static __forceinline__ ival det(const ival (&m)[3][3]) {
        ival t0 = imul_nl(m[0][0], det2(m[1][1], m[1][2], m[2][1], m[2][2]));
        ival t1 = imul_nl(m[0][1], det2(m[1][0], m[1][2], m[2][0], m[2][2]));
        ival t2 = imul_nl(m[0][2], det2(m[1][0], m[1][1], m[2][0], m[2][1]));
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
        ival acc = imul_nl(F[0][0], minor_det<4>(F, 0, 0));
#pragma unroll
        for (int j = 1; j < 4; j++) {
            ival term = imul_nl(F[0][j], minor_det<4>(F, 0, j));
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
            diag = iadd(diag, imul_nl(cof, Ad[j][i]));
        }
        tr = iadd(tr, diag);
    }
    return tr;
}

// decision of topFrame::countZeros for one sub-interval: 0 = certified no zero, 1 = certified zero, 2 = failure
__device__ __forceinline__ int count_decide(const ival& time, const ival& L, const ival& R, const ival& V, const ival& Dv) {
    ival LR = imul_nl(L, R);
    if (LR.lo > 0) {
        if (idefinite(V)) return 0;
        if (idefinite(Dv)) return 0;
        ival w = isub(time, pt(time.lo));
        ival mvt = imul_nl(w, Dv);
        ival half = mk(mvt.lo * 0.5, mvt.hi * 0.5);
        ival bound = ihull(iadd(L, half), isub(R, half));
        return idefinite(bound) ? 0 : 2;
    } else if (LR.hi < 0) {
        return idefinite(Dv) ? 1 : 2;
    }
    return 2;
}

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
#pragma unroll 4
    for (int k = p - 1; k >= 0; k--) acc = horner_pt(acc, tau, mt2iv(s[k]));
    return acc;
}
// sum_{k>=1} s_k tau^k by Horner: the INCREMENT of a series over a step.  Leaving the order-0 coefficient out keeps the rounding of
// the increment on its own, much finer, grid (oracle/c2.hpp:eval_inc): the centres x, Tc advance by x+ = x + mid(inc) and the
// residual (x - x+) + inc is thin, instead of carrying ~1 ulp(x) of noise per step that the unstable directions amplify by e^{0.7 t}.
__device__ __forceinline__ ival eval_inc(const mt* s, int p, double tau) {
    ival acc = mt2iv(s[p]);
#pragma unroll 4
    for (int k = p - 1; k >= 1; k--) acc = horner_pt(acc, tau, mt2iv(s[k]));
    return mk(__dmul_rd(acc.lo, tau), __dmul_ru(acc.hi, tau));
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


// ---- helpers of the second-variation / set-update phases (oracle/c2.hpp, same expression order) ----
__device__ __forceinline__ ival ihalf(const ival& a) { return mk(0.5 * a.lo, 0.5 * a.hi); }
__device__ __forceinline__ ival idivk_tab(const ival& a, const double* ilo, const double* ihi, int k);
// a / k directly in (mid,t) form (oracle/c1.hpp:scalek): no interval conversion, no sign selects on the critical chain of an order.
//   m' = RN(m ih),  t' = RU(t ihs + 2^-1000)  >=  |m'| + (t - |m|) ih + |m| (ih - 1/k) + rounding of m'     (ih = RU(1/k), ihs = RU(ih (1 + 2^-50)))
__device__ __forceinline__ mt mt_scale(const mt& a, double ih, double ihs) {
    mt o; o.m = a.m * ih; o.t = __fma_ru(a.t, ihs, 0x1p-1000);
    return o;
}
__device__ __forceinline__ mt ptm(double v) { mt o; o.m = v; o.t = fabs(v); return o; }      // iv2mt of a point
// a += x * p for a point p: one directed FMA per bound (tighter than a rounded product plus a rounded sum)
__device__ __forceinline__ void ipfma(ival& a, const ival& x, double p) {
    const double l = p >= 0 ? x.lo : x.hi, h = p >= 0 ? x.hi : x.lo;
    a.lo = __fma_rd(l, p, a.lo); a.hi = __fma_ru(h, p, a.hi);
}

// host-side twin of oracle/common.hpp:step_count
__host__ __device__ inline int step_count(double L_minus, double h) {
    int S = (int)ceil(L_minus / h);
    while ((double)(S - 1) * h >= L_minus) S--;
    while ((double)S * h < L_minus) S++;
    return S;
}

// ------------------------------------------------------------------------------------------------
// Cauchy tasks.  Every Taylor-table product of one order is one "task" = sum_j a(j) b(kk-j) with
//   a(0) = T[a0r][a0c], a(j) = T[iA][j];   b(0) = T[b0r][b0c], b(m) = T[iB][m]     (z = 1 - u has its own rows).
// A CTA of NT = 64 threads owns one front: warp "phi" holds the A tasks (g = u z, q = w w), the B tasks (H w) and the lanes that turn
// g, q into the M-series and H; warp "Psi" holds the C tasks (M Psi^u), in pairs that share the M-series (or, CCP_SINGLES, one per lane
// with two columns of Psi hosted by idle lanes of the phi warp).  table_loop<N> runs them in stages of two Taylor orders; the sums keep
// the summation order of oracle/c1.hpp, which computes them serially.  Which lane does what is a table (LaneC, make_lanec).
// ------------------------------------------------------------------------------------------------
constexpr int NT = 64;                  // threads per front (one CTA): warp 0 = front tables (phi), warp 1 = variational tables (Psi)
constexpr int NWARP = NT / 32;

struct Task { short iA, iB, a0r, a0c, b0r, b0c, out; signed char kind, flip; };   // kind: -1 none, 0 A, 1 B, 2 C

// kind 0 (A): lane l < NA of the phi warp;  kind 1 (B): lane l < NB of the phi warp;
// kind 2 (C): C task index cidx = round*32 + lane of the Psi warp.
template <int N>
__device__ inline Task decode_task(int kind, int idx) {
    using L = Lay<N>;
    Task t; t.kind = -1; t.iA = t.iB = t.a0r = t.a0c = t.b0r = t.b0c = t.out = 0; t.flip = 0;
    if (kind == 0) {
        if (idx >= L::NA) return t;
        t.kind = 0;
        if (idx < N) {                         // g[i] = u*z
            const int i = idx;
            t.iA = L::U(i); t.iB = L::Z(i);
            t.a0r = L::U(i); t.a0c = 0; t.b0r = L::Z(i); t.b0c = 0;
            t.out = L::G(i);
        } else {                                // q[e] = w_e*w_{e+1}
            const int e = idx - N;
            t.iA = L::U(e); t.iB = L::U(e + 1);
            t.a0r = L::G(e); t.a0c = KMAX - 1; t.b0r = L::G(e + 1); t.b0c = KMAX - 1;
            t.out = L::Q(e);
        }
        return t;
    }
    if (kind == 1) {
        if (idx >= L::NB) return t;
        int c = 0;
        for (int i = 0; i < N; i++) {
            if (c == idx) {
                t.kind = 1;
                t.iA = L::G(i); t.iB = L::U(i);
                t.a0r = L::G(i); t.a0c = 0; t.b0r = L::G(i); t.b0c = KMAX - 1;
                t.out = idx;
            }
            c++;
        }
        return t;
    }
    if (idx < L::NC) {
        int c = 0;
        for (int i = 0; i < N; i++) for (int cc = 0; cc < 2 * N; cc++) for (int w = 0; w < 3; w++) {
            if ((w == 1 && i == 0) || (w == 2 && i == N - 1)) continue;
            if (c == idx) {
                t.kind = 2;
                t.iA = (w == 0) ? L::MD(i) : (w == 1 ? L::MO(i - 1) : L::MO(i));
                t.iB = (w == 0) ? L::PU(i, cc) : (w == 1 ? L::PU(i - 1, cc) : L::PU(i + 1, cc));
                t.a0r = t.iA; t.a0c = 0; t.b0r = t.iB; t.b0c = 0;
                t.out = L::NB + idx;
            }
            c++;
        }
    }
    return t;
}
// C tasks are processed in PAIRS that share their first operand (the same M-series times two columns of Psi):
// pair P -> group (i,w) = P / (D/2) in the enumeration order of decode_task, columns 2*(P % (D/2)) and +1.
template <int N>
__device__ inline void pair_tasks(int P, int& idx0, int& idx1) {
    constexpr int D = 2 * N, H = D / 2;
    idx0 = idx1 = 1 << 20;
    if (P < 0 || P >= Lay<N>::NC / 2) return;
    const int grp = P / H, pr = P % H;
    int gcount = 0;
    for (int i = 0; i < N; i++) {
        const int nw = 1 + (i > 0 ? 1 : 0) + (i < N - 1 ? 1 : 0);
        int widx = 0;
        for (int w = 0; w < 3; w++) {
            if ((w == 1 && i == 0) || (w == 2 && i == N - 1)) continue;
            if (gcount == grp) {
                const int base = (i == 0 ? 0 : (3 * i - 1) * D);
                idx0 = base + (2 * pr) * nw + widx;
                idx1 = base + (2 * pr + 1) * nw + widx;
            }
            gcount++; widx++;
        }
    }
}
// Lane -> pair assignment of the Psi warp.  A quarter warp (8 lanes) reads one 16-byte coefficient per lane with one LDS.128; rows r and r'
// collide when r = r' (mod 8) (row stride 84 words).  The Psi^u rows a pair reads are 16 + 6 i' + 2 pr (+1): only the residues {0,2,4,6}
// (resp. {1,3,5,7}) occur, so a conflict-free quarter warp holds at most four distinct rows, one per residue.  For n = 3 such an assignment of
// the 21 pairs to three quarter warps exists (found by hand; the M-series rows 11..15 never collide); the enumeration order would put rows with
// equal residues side by side (two-way conflicts on two of the three loads of every term of the pass).
template <int N> __device__ inline int psi_lane_to_pair(int lane) { return lane; }
template <> __device__ inline int psi_lane_to_pair<3>(int lane) {
    constexpr signed char perm[21] = { 0, 9, 1, 10, 2, 11, 3, 6,   4, 7, 19, 5, 12, 15, 13, 16,   14, 17, 18, 8, 20 };
    return lane < 21 ? perm[lane] : lane;
}
// Step image: which two of the NH = D + D^2 series (h < D: u, v;  h >= D: Psi entry h - D, table rows 16 + h - D for n = 3) a lane evaluates.
// Default: neighbours 2 tt, 2 tt + 1 -- the lanes of a quarter warp then read every second row (two- and three-way bank conflicts).  n = 3:
// rows with distinct residues mod 8 in every quarter warp for both loads (same argument as psi_lane_to_pair).
template <int N> __device__ __forceinline__ int image_series(int tt, int which) {
    constexpr int NH = 2 * N + 4 * N * N;
    const int h = 2 * tt + which;
    return h < NH ? h : NH;
}
__constant__ signed char c_image3[2][24] = {       // in constant memory: a table local to the function would be rebuilt on the stack at every call
    { 0, 1, 2, 3, 4, 5, 12, 13,   6, 7, 8, 9, 10, 11, 20, 21,   14, 15, 16, 17, 18, 42, 42, 42 },
    { 22, 23, 24, 25, 26, 27, 28, 29,   30, 31, 32, 33, 34, 35, 36, 37,   38, 39, 40, 41, 19, 42, 42, 42 } };
template <> __device__ __forceinline__ int image_series<3>(int tt, int which) { return c_image3[which][tt]; }
template <int N> struct PhaseCheck {
    using L = Lay<N>;
    static_assert(L::NA + L::NB <= 32, "the phi warp holds the A tasks and the B tasks side by side");
    static constexpr int CROUNDS = (L::NC / 2 + 31) / 32;      // rounds of C-task PAIRS in the Psi warp
    static_assert(CROUNDS <= 2 && L::NC % 2 == 0, "Psi warp: at most 2 rounds of C-task pairs");
    static_assert(N * 2 * N <= 32, "Psi warp: one combine task per lane");
};

__device__ __forceinline__ void bar_pair() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

// Per-thread task registers.  The fields are laundered through an empty asm so that the compiler keeps them
// in registers instead of re-deriving them from threadIdx with select chains inside the hot loop.
// Table offsets are < 2^11 (rows*21), scratch slots < 2^8: three words per task keep the loop-invariant state of four tasks
// plus the combine operands in registers (unpacked with one shift/mask per use; a spilled word costs an LDL round trip).
struct TaskR {
    unsigned ab, zz, ok;     // ab = offA | offB<<16 ; zz = offa0 | offb0<<16 ; ok = out | (kind+1)<<16
    __device__ __forceinline__ int offA() const { return (int)(ab & 0xffffu); }
    __device__ __forceinline__ int offB() const { return (int)(ab >> 16); }
    __device__ __forceinline__ int offa0() const { return (int)(zz & 0xffffu); }
    __device__ __forceinline__ int offb0() const { return (int)(zz >> 16); }
    __device__ __forceinline__ int out() const { return (int)(ok & 0xffffu); }
    __device__ __forceinline__ int kind() const { return (int)(ok >> 16) - 1; }
};
__device__ __forceinline__ int opaque(int v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ unsigned opaqueu(unsigned v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ TaskR make_taskr(const Task& t) {
    TaskR r;
    r.ab = opaqueu((unsigned)((int)t.iA * KMAX) | ((unsigned)((int)t.iB * KMAX) << 16));
    r.zz = opaqueu((unsigned)((int)t.a0r * KMAX + t.a0c) | ((unsigned)((int)t.b0r * KMAX + t.b0c) << 16));
    r.ok = opaqueu((unsigned)t.out | ((unsigned)(t.kind + 1) << 16));
    return r;
}
// Cauchy sums of the orders J (s0) and J+1 (s1), J even, of one task in one pass over the operands.  Summation order of every sum
// (oracle/c1.hpp:cauchy_order): interior terms l = 1..order-1 ascending, then the end terms l = 0 and l = order:
//   order J   :  sum_{l=1}^{J-1} a[l] b[J-l]   +  a0 b[J]    +  a[J] b0
//   order J+1 :  sum_{l=1}^{J}   a[l] b[J+1-l] +  a0 b[J+1]  +  a[J+1] b0
// a(0), b(0) live at their own offsets (w0 = u0 - 1/2 and z0 = 1 - u0 differ from u only in order 0).
// pass_early: everything that does not involve a[J], a[J+1] (for the products H*w and M*Psi^u these are the coefficients the stage is
// waiting for) and precedes them in the summation order: 2 LDS.128 + 8 DFMA per term, 8 independent FMA chains.
// MODE 1 ("FAST") = the stage is known to have J >= 2 and a second order, MODE 2 = J == 0 and a second order (every stage of an even-order
// run is one of the two; MODE 0 = anything, decided at run time): no branches, so that
// the two orders' dependent chains end up in one basic block and the scheduler interleaves them (a warp issues in order).
template <int MODE>
__device__ __forceinline__ void pass_early(const mt* __restrict__ Tf, const TaskR& t, int J, acc4& s0, acc4& s1) {
    if (MODE == 1 || (MODE == 0 && J >= 2)) {
        const mt* pa = Tf + t.offA() + 1;                        // a[l]
        const mt* pb = Tf + t.offB() + J - 1;                    // b[J-l]
        mt yp = pb[1];                                           // b[J+1-l]: the only operand carried from term to term
        int l = 1;
#pragma unroll UNROLL_S
        for (; l + 1 < J; l += 2) {                              // two terms per iteration: no operand rotation through register moves
            const mt x0 = pa[0], x1 = pa[1], y0 = pb[0], y1 = pb[-1];
            acc_term(s1, x0, yp); acc_term(s0, x0, y0); acc_term(s1, x1, y0); acc_term(s0, x1, y1);
            yp = y1; pa += 2; pb -= 2;
        }
        { const mt x0 = pa[0], y0 = pb[0]; acc_term(s1, x0, yp); acc_term(s0, x0, y0); }      // J - 1 is odd: one term is left
        acc_term(s0, Tf[t.offa0()], Tf[t.offB() + J]);
    }
}
// pass_late: the remaining terms, in order.  two = the stage has a second order.
template <int MODE>
__device__ __forceinline__ void pass_late(const mt* __restrict__ Tf, const TaskR& t, int J, bool two, acc4& s0, acc4& s1) {
    const mt a0 = Tf[t.offa0()], b0 = Tf[t.offb0()];
    if (MODE == 1 || (MODE == 0 && J >= 2)) {
        const mt aJ = Tf[t.offA() + J];
        acc_term(s0, aJ, b0);
        if (MODE == 1 || two) {
            acc_term(s1, aJ, Tf[t.offB() + 1]);
            acc_term(s1, a0, Tf[t.offB() + J + 1]);
            acc_term(s1, Tf[t.offA() + J + 1], b0);
        }
    } else {
        acc_term(s0, a0, b0);
        if (MODE == 2 || two) { acc_term(s1, a0, Tf[t.offB() + 1]); acc_term(s1, Tf[t.offA() + 1], b0); }
    }
}
// Two tasks that SHARE their first operand series (Psi warp: one M-series times two columns of Psi): 3 LDS.128 + 16 DFMA per term.
template <int MODE>
__device__ __forceinline__ void pair_early(const mt* __restrict__ Tf, const TaskR& t0, const TaskR& t1, int J,
                                           acc4& a0, acc4& a1, acc4& b0, acc4& b1) {
    if (MODE == 1 || (MODE == 0 && J >= 2)) {
        const mt* pa = Tf + t0.offA() + 1; const mt* pb = Tf + t0.offB() + J - 1; const mt* qb = Tf + t1.offB() + J - 1;
        mt yp = pb[1], zp = qb[1];
        int l = 1;
#pragma unroll UNROLL_P
        for (; l + 1 < J; l += 2) {
            const mt x0 = pa[0], x1 = pa[1], y0 = pb[0], y1 = pb[-1], z0 = qb[0], z1 = qb[-1];
            acc_term(a1, x0, yp); acc_term(b1, x0, zp); acc_term(a0, x0, y0); acc_term(b0, x0, z0);
            acc_term(a1, x1, y0); acc_term(b1, x1, z0); acc_term(a0, x1, y1); acc_term(b0, x1, z1);
            yp = y1; zp = z1; pa += 2; pb -= 2; qb -= 2;
        }
        { const mt x0 = pa[0], y0 = pb[0], z0 = qb[0]; acc_term(a1, x0, yp); acc_term(b1, x0, zp); acc_term(a0, x0, y0); acc_term(b0, x0, z0); }
        const mt x0e = Tf[t0.offa0()];
        acc_term(a0, x0e, Tf[t0.offB() + J]); acc_term(b0, x0e, Tf[t1.offB() + J]);
    }
}
template <int MODE>
__device__ __forceinline__ void pair_late(const mt* __restrict__ Tf, const TaskR& t0, const TaskR& t1, int J, bool two,
                                          acc4& a0, acc4& a1, acc4& b0, acc4& b1) {
    const mt x0 = Tf[t0.offa0()], y0 = Tf[t0.offb0()], z0 = Tf[t1.offb0()];
    if (MODE == 1 || (MODE == 0 && J >= 2)) {
        const mt xJ = Tf[t0.offA() + J];
        acc_term(a0, xJ, y0); acc_term(b0, xJ, z0);
        if (MODE == 1 || two) {
            acc_term(a1, xJ, Tf[t0.offB() + 1]);     acc_term(b1, xJ, Tf[t1.offB() + 1]);
            acc_term(a1, x0, Tf[t0.offB() + J + 1]); acc_term(b1, x0, Tf[t1.offB() + J + 1]);
            const mt xJ1 = Tf[t0.offA() + J + 1];
            acc_term(a1, xJ1, y0); acc_term(b1, xJ1, z0);
        }
    } else {
        acc_term(a0, x0, y0); acc_term(b0, x0, z0);
        if (MODE == 2 || two) {
            acc_term(a1, x0, Tf[t0.offB() + 1]); acc_term(b1, x0, Tf[t1.offB() + 1]);
            const mt x1 = Tf[t0.offA() + 1];
            acc_term(a1, x1, y0); acc_term(b1, x1, z0);
        }
    }
}
// product with a parameter interval of known sign: c = sg * [pl, ph], 0 <= pl <= ph.  Same end points as imul_nl(a, c).
__device__ __forceinline__ SPar make_spar(const ival& c) {
    SPar s; s.straddle = (c.lo < 0 && c.hi > 0) ? 1 : 0; s.neg = (c.hi <= 0 && c.lo < 0) ? 1 : 0;
    s.pl = s.neg ? -c.hi : c.lo; s.ph = s.neg ? -c.lo : c.hi;
    return s;
}
__device__ __forceinline__ ival imul_pos3(const ival& a, double pl, double ph, int neg) {   // a * (neg ? -[pl,ph] : [pl,ph]), 0 <= pl <= ph
    const double lo = __dmul_rd(a.lo, a.lo >= 0 ? pl : ph), hi = __dmul_ru(a.hi, a.hi >= 0 ? ph : pl);
    return neg ? mk(-hi, -lo) : mk(lo, hi);
}
__device__ __forceinline__ ival imul_par(const ival& a, const SPar& s, const ival& c) {
    if (s.straddle) return imul_nl(a, c);
    const double lo = __dmul_rd(a.lo, a.lo >= 0 ? s.pl : s.ph), hi = __dmul_ru(a.hi, a.hi >= 0 ? s.ph : s.pl);
    return s.neg ? mk(-hi, -lo) : mk(lo, hi);
}
// a / k as a product with [rd(1/k), ru(1/k)]  (oracle/c1.hpp:divk)
__device__ __forceinline__ ival idivk_tab(const ival& a, const double* ilo, const double* ihi, int k) {
    return mk(__dmul_rd(a.lo, a.lo >= 0 ? ilo[k] : ihi[k]), __dmul_ru(a.hi, a.hi >= 0 ? ihi[k] : ilo[k]));
}


// oracle/c2.hpp:c2_majorant -- scalar majorant of the second variation xi' = Df xi + D^2f[Psi c, psi] on the a-priori box, per unit |c|_1:
//   |Xi(tau)| <= g1 tau,   |Xi(tau) - sum_{k<=4} Xi_k tau^k| <= g5 tau^5      (out[0] = g1, out[1] = g5)
template <int N>
__device__ inline void majorant2(const ival* b, const ival* c, double h, double* out) {
    double bh = 0, cs = 0;
    for (int i = 0; i < N; i++) bh = dmax(bh, imag(b[i]));
    for (int i = 0; i < N; i++) {
        double s = 0;
        if (i > 0) s = __dadd_ru(s, imag(c[i - 1]));
        if (i < N - 1) s = __dadd_ru(s, imag(c[i]));
        cs = dmax(cs, s);
    }
    const double W0 = 0.5625, G0 = 0.25, V0 = 0.5;
    constexpr int K = 5;
    double bc = __dadd_ru(bh, cs);
    double F0 = __dmul_ru(__dmul_ru(bc, G0), W0);
    double w[K + 2], g[K + 2], be[K + 2], Mh[K + 2], ga[K + 2], bb[K + 2];
    w[0] = W0; g[0] = G0; w[1] = dmax(V0, F0);
    for (int k = 1; k <= K; k++) {
        double s = 0; for (int j = 0; j <= k; j++) s = __fma_ru(w[j], w[k - j], s);
        g[k] = s;
        double f = 0; for (int j = 0; j <= k; j++) f = __fma_ru(g[j], w[k - j], f);
        f = __dmul_ru(bc, f);
        w[k + 1] = __ddiv_ru(dmax(w[k], f), (double)(k + 1));
    }
    double M0 = __dadd_ru(__dadd_ru(__dmul_ru(bh, 0.75), __dmul_ru(cs, 0.25)), __dmul_ru(__dmul_ru(2.0, cs), __dmul_ru(W0, W0)));
    double a = dmax(1.0, M0), ah = __dmul_ru(a, h);
    be[0] = __dadd_ru(__dadd_ru(1.0, ah), __dmul_ru(ah, ah));
    Mh[0] = M0;
    double k3 = __dmul_ru(3.0, bc);
    for (int k = 1; k <= K; k++) Mh[k] = __dmul_ru(k3, g[k]);
    for (int k = 0; k <= K; k++) {
        double s = 0; for (int j = 0; j <= k; j++) s = __fma_ru(Mh[j], be[k - j], s);
        be[k + 1] = __ddiv_ru(dmax(be[k], s), (double)(k + 1));
    }
    const double H6 = __dmul_ru(6.0, bc);
    ga[0] = __dmul_ru(__dmul_ru(h, __dmul_ru(be[0], __dmul_ru(be[0], be[0]))), __dmul_ru(H6, W0));
    for (int k = 0; k <= K; k++) { double s = 0; for (int j = 0; j <= k; j++) s = __fma_ru(be[j], be[k - j], s); bb[k] = s; }
    for (int k = 0; k < K; k++) {
        double s = 0; for (int j = 0; j <= k; j++) s = __fma_ru(Mh[j], ga[k - j], s);
        double f = 0; for (int j = 0; j <= k; j++) f = __fma_ru(w[j], bb[k - j], f);
        s = __fma_ru(H6, f, s);
        ga[k + 1] = __ddiv_ru(dmax(ga[k], s), (double)(k + 1));
    }
    out[0] = ga[1]; out[1] = ga[5];
}

// oracle/c2.hpp:frame_of_state, one entry:  S0 = sum_j Theta_j r0_j ;  W = ((Tc + S0) + R) + Ps Err
template <int N>
__device__ __forceinline__ ival frame_entry(const FrontSmem<N>& S, int cc, int bb, ival* s0out) {
    constexpr int D = 2 * N;
    ival a = pt(0.0);
#pragma unroll
    for (int j = 0; j < D; j++) ipfma(a, S.r0[j], S.Th[j][cc][bb]);
    *s0out = a;
    acc4 e = acc_zero();
#pragma unroll
    for (int j = 0; j < N; j++) acc_term(e, iv2mt(S.Ps[cc][j]), S.Errm[j][bb]);
    return iadd(iadd(iadd(pt(S.Tc[cc][bb]), a), S.R[cc][bb]), acc_finish(e));
}

// ---- second variation over the hull (oracle/c2.hpp): directions c_j = C e_j (j < D, weight r0_j) and the box r (j = D, weight 1).
//      All sums are dense fixed-length dot products in (mid,t) form (missing chain neighbours = zero parameters), so every stage is
//      one branch-free instruction stream over (direction, entry) items.  Called by all threads of the CTA. ----
template <int N>
__device__ __noinline__ void second_variation(FrontSmem<N>& S, PostTab<N>& PT, int cur, double hs, double g5, bool frames) {
    constexpr int D = 2 * N;
    const int tid = threadIdx.x ^ S.vswap;
    HullM<N>& H = S.hm;
    // H (warp 0 only; warp 1 waits at the barrier): lanes < N: W_0, W_1, G = u(1-u), Pw = w_0 w_1; lanes N..2N-2: off-diagonals of
    //    M_0, M_1; then, after a __syncwarp, lanes < N: u_2 = F(u)/2 and the diagonals of M_0, M_1 from G, Pw of the chain neighbours
    if (tid < 32) {
        if (tid < N) {
            const int m = tid;
            const ival u = S.X[m];
            const mt w0 = iv2mt(isub(u, pt(0.5))), w1 = iv2mt(S.X[N + m]);
            H.W[0][m] = w0; H.W[1][m] = w1; H.G[m] = iv2mt(imul(u, isub(pt(1.0), u)));
            acc4 a = acc_zero(); acc_term(a, w0, w1); H.Pw[m] = iv2mt(acc_finish(a));
        } else if (tid < 2 * N - 1) {
            const int e = tid - N;
            const mt w0a = iv2mt(isub(S.X[e], pt(0.5))), w0b = iv2mt(isub(S.X[e + 1], pt(0.5))), w1a = iv2mt(S.X[N + e]), w1b = iv2mt(S.X[N + e + 1]);
            acc4 a = acc_zero(); acc_term(a, w0a, w0b);
            acc4 a2 = acc_zero(); acc_term(a2, S.pm.C2[e], iv2mt(acc_finish(a)));
            const mt m0 = iv2mt(acc_finish(a2));
            H.M0[e][e + 1] = m0; H.M0[e + 1][e] = m0;
            acc4 u = acc_zero(); acc_term(u, w0a, w1b); acc_term(u, w1a, w0b);
            acc4 u2 = acc_zero(); acc_term(u2, S.pm.C2[e], iv2mt(acc_finish(u)));
            const mt m1 = iv2mt(acc_finish(u2));
            H.M1[e][e + 1] = m1; H.M1[e + 1][e] = m1;
        }
        __syncwarp();
        if (tid < N) {
            const int i = tid, im = i > 0 ? i - 1 : 0, ip = i < N - 1 ? i + 1 : N - 1;
            const mt gi = H.G[i], gm = H.G[im], gp = H.G[ip];
            acc4 sa = acc_zero(); acc_term(sa, S.pm.B[i], gi); acc_term(sa, S.pm.Cm[i], gm); acc_term(sa, S.pm.Cp[i], gp);
            H.W[2][i] = iv2mt(ihalf(ineg(imul(isub(S.X[i], pt(0.5)), acc_finish(sa)))));
            acc4 ta = acc_zero(); acc_term(ta, S.pm.B3[i], gi); acc_term(ta, S.pm.Cm[i], gm); acc_term(ta, S.pm.Cp[i], gp);
            H.M0[i][i] = iv2mt(isub(S.pm.halfB[i], acc_finish(ta)));
            acc4 ma = acc_zero(); acc_term(ma, S.pm.B6[i], H.Pw[i]); acc_term(ma, S.pm.C2m[i], H.Pw[im]); acc_term(ma, S.pm.C2p[i], H.Pw[ip]);
            H.M1[i][i] = iv2mt(acc_finish(ma));
        }
    }
    __syncthreads();
    // X1: u_3 = M_0 u_1 / 6;  d_0..d_3 of every direction
    if (tid >= 32 && tid < 32 + N) {
        const int i = tid - 32;
        acc4 a = acc_zero();
#pragma unroll
        for (int l = 0; l < N; l++) acc_term(a, H.M0[i][l], H.W[1][l]);
        H.W[3][i] = iv2mt(idivk_tab(acc_finish(a), c_invlo, c_invhi, 6));
    }
    for (int t = tid; t < (D + 1) * N; t += NT) {
        const int j = t / N, m = t % N;
        mt d0[N], d1[N];
#pragma unroll
        for (int l = 0; l < N; l++) {
            if (j < D) { d0[l] = ptm(S.C[cur][l][j]); d1[l] = ptm(S.C[cur][N + l][j]); }
            else       { d0[l] = iv2mt(S.r[cur][l]);  d1[l] = iv2mt(S.r[cur][N + l]); }
        }
        acc4 a2 = acc_zero(), a3 = acc_zero();
#pragma unroll
        for (int l = 0; l < N; l++) acc_term(a2, H.M0[m][l], d0[l]);
#pragma unroll
        for (int l = 0; l < N; l++) acc_term(a3, H.M1[m][l], d0[l]);
#pragma unroll
        for (int l = 0; l < N; l++) acc_term(a3, H.M0[m][l], d1[l]);
        mt dm0 = d0[0], dm1 = d1[0];
#pragma unroll
        for (int l = 1; l < N; l++) if (l == m) { dm0 = d0[l]; dm1 = d1[l]; }
        PT.u.x.DD[j][0][m] = dm0; PT.u.x.DD[j][1][m] = dm1;
        PT.u.x.DD[j][2][m] = iv2mt(ihalf(acc_finish(a2)));
        PT.u.x.DD[j][3][m] = iv2mt(idivk_tab(acc_finish(a3), c_invlo, c_invhi, 6));
    }
    __syncthreads();
    // X2a: product series (w_i d_m)_k, k <= 3, all N x N pairs.  One thread per (direction j, i, m): the four orders are independent
    //      chains over the same eight operands (loaded once); same sums in the same order as an item per (j, k, i, m)
    constexpr int NJM = (D + 1) * N * N;
    for (int t = tid; t < NJM; t += NT) {
        const int j = t / (N * N), i = (t / N) % N, m = t % N;
        const mt w0 = H.W[0][i], w1 = H.W[1][i], w2 = H.W[2][i], w3 = H.W[3][i];
        const mt e0 = PT.u.x.DD[j][0][m], e1 = PT.u.x.DD[j][1][m], e2 = PT.u.x.DD[j][2][m], e3 = PT.u.x.DD[j][3][m];
        acc4 a0 = acc_zero(), a1 = acc_zero(), a2 = acc_zero(), a3 = acc_zero();
        acc_term(a0, w0, e0);
        acc_term(a1, w0, e1); acc_term(a1, w1, e0);
        acc_term(a2, w0, e2); acc_term(a2, w1, e1); acc_term(a2, w2, e0);
        acc_term(a3, w0, e3); acc_term(a3, w1, e2); acc_term(a3, w2, e1); acc_term(a3, w3, e0);
        PT.u.x.WD[j][0][i][m] = iv2mt(acc_finish(a0)); PT.u.x.WD[j][1][i][m] = iv2mt(acc_finish(a1));
        PT.u.x.WD[j][2][i][m] = iv2mt(acc_finish(a2)); PT.u.x.WD[j][3][i][m] = iv2mt(acc_finish(a3));
    }
    __syncthreads();
    // X2b: Taylor coefficients N_k of DM(u)[Psi^u c_j] (dense): three parameter * product-series slots per entry; one thread per (j, i, l),
    //      the four orders share the parameters and the slot indices
    for (int t = tid; t < NJM; t += NT) {
        const int j = t / (N * N), i = (t / N) % N, l = t % N;
        // slots: diagonal (i,i): 6 b_i (w_i d_i), 2 c_{i-1} (w_{i-1} d_{i-1}), 2 c_i (w_{i+1} d_{i+1});  off-diagonal e = min(i,l): 2 c_e (w_{e+1} d_e + w_e d_{e+1})
        const int e = i < l ? i : l, e1 = e + 1 < N ? e + 1 : N - 1;
        const int a0 = i == l ? i : e1, b0 = i == l ? i : e;
        const int a1 = i == l ? (i > 0 ? i - 1 : 0) : e, b1 = i == l ? a1 : e1;
        const int a2 = i == l ? (i < N - 1 ? i + 1 : N - 1) : e, b2 = i == l ? a2 : e1;
        const mt p0 = S.pm.PN[i][l][0], p1 = S.pm.PN[i][l][1], p2 = S.pm.PN[i][l][2];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const mt (*WDk)[N] = PT.u.x.WD[j][k];
            acc4 a = acc_zero();
            acc_term(a, p0, WDk[a0][b0]);
            acc_term(a, p1, WDk[a1][b1]);
            acc_term(a, p2, WDk[a2][b2]);
            PT.u.x.NN[j][k][i][l] = iv2mt(acc_finish(a));      // NN overlays DD (dead since X2a)
        }
    }
    __syncthreads();
    // X3: the four n x n blocks of Xi_j(hs) = sum_{k<=4} Xi_k hs^k + remainder, one (j, i, k) entry of each block per item
    const double hs2 = __dmul_ru(hs, hs), hs5 = __dmul_ru(__dmul_ru(hs2, hs2), hs);
    for (int t = tid; t < (D + 1) * N * N; t += NT) {
        const int j = t / (N * N), i = (t / N) % N, k = t % N;
        double c1n = 0;
#pragma unroll
        for (int q = 0; q < D; q++) c1n = __dadd_ru(c1n, j < D ? fabs(S.C[cur][q][j]) : imag(S.r[cur][q]));
        const double rem = __dmul_ru(__dmul_ru(g5, hs5), c1n);
        const mt (*NN)[N][N] = PT.u.x.NN[j];
        acc4 p00 = acc_zero(), p01 = acc_zero(), p10 = acc_zero();
#pragma unroll
        for (int l = 0; l < N; l++) acc_term(p00, H.M0[i][l], NN[0][l][k]);
#pragma unroll
        for (int l = 0; l < N; l++) acc_term(p00, NN[0][i][l], H.M0[l][k]);
#pragma unroll
        for (int l = 0; l < N; l++) acc_term(p01, H.M0[i][l], NN[1][l][k]);
#pragma unroll
        for (int l = 0; l < N; l++) acc_term(p01, NN[0][i][l], H.M1[l][k]);
#pragma unroll
        for (int l = 0; l < N; l++) acc_term(p10, H.M1[i][l], NN[0][l][k]);
#pragma unroll
        for (int l = 0; l < N; l++) acc_term(p10, NN[1][i][l], H.M0[l][k]);
        const ival P00 = acc_finish(p00), P01 = acc_finish(p01), P10 = acc_finish(p10);
        const ival N0 = mt2iv(NN[0][i][k]), N1 = mt2iv(NN[1][i][k]), N2 = mt2iv(NN[2][i][k]), N3 = mt2iv(NN[3][i][k]);
        const double* il = c_invlo; const double* ih = c_invhi;
        const ival X3 = idivk_tab(iadd(N2, ihalf(P00)), il, ih, 3);
        const ival Z = pt(0.0);
        const ival hN0 = ihalf(N0);
        // cf[block][order-1]: blocks 0 UU, 1 UV, 2 VU, 3 VV
        const ival c3[4] = { idivk_tab(X3, il, ih, 4), idivk_tab(N1, il, ih, 12),
                             idivk_tab(iadd(iadd(N3, idivk_tab(P01, il, ih, 6)), ihalf(P10)), il, ih, 4), idivk_tab(iadd(N2, idivk_tab(P00, il, ih, 6)), il, ih, 4) };
        const ival c2[4] = { idivk_tab(N1, il, ih, 6), idivk_tab(N0, il, ih, 6), X3, idivk_tab(N1, il, ih, 3) };
        const ival c1[4] = { hN0, Z, ihalf(N1), hN0 };
        const ival c0[4] = { Z, Z, N0, Z };
#pragma unroll
        for (int q = 0; q < 4; q++) {
            ival acc = c3[q]; double am = imag(c3[q]);
            acc = horner_pt(acc, hs, c2[q]); am = __fma_ru(am, hs, imag(c2[q]));
            acc = horner_pt(acc, hs, c1[q]); am = __fma_ru(am, hs, imag(c1[q]));
            acc = horner_pt(acc, hs, c0[q]); am = __fma_ru(am, hs, imag(c0[q]));
            acc = horner_pt(acc, hs, Z); am = __dmul_ru(am, hs);
            const int row = (q >> 1) * N + i, col = (q & 1) * N + k;
            PT.Xi[j][row][col] = iadd(acc, sym(rem));
            PT.u.x.AM[j][row][col] = __dadd_ru(am, rem);        // AM overlays WD (dead since X2b)
        }
    }
    __syncthreads();
    // X4: dJ = sum_j Xi_j r0_j + Xi_r encloses DPhi(phi) - DPhi(x) on the set; J = Jc + [0,1] dJ for the mean-value form of the
    //     phi-set update; magX >= sup over [0,hs] of |Psi(tau;phi) - Psi(tau;x)|
    for (int t = tid; t < D * D; t += NT) {
        const int row = t / D, col = t % D;
        acc4 a = acc_zero(); double mg = 0.0;
#pragma unroll
        for (int j = 0; j <= D; j++) {
            acc_term(a, iv2mt(PT.Xi[j][row][col]), S.r0m[j]);
            mg = __fma_ru(PT.u.x.AM[j][row][col], j < D ? imag(S.r0[j]) : 1.0, mg);
        }
        const ival dj = acc_finish(a);
        PT.dJm[row][col] = iv2mt(dj);
        const ival jc = S.Jc[row][col], jf = iadd(jc, ihull(dj, pt(0.0)));
        S.J[row][col] = jf; PT.Jm[row][col] = iv2mt(jf); PT.Jcm[row][col] = iv2mt(jc);
        if (row < N) PT.magXu[row][col] = mg; else PT.magXv[row - N][col] = mg;
    }
    __syncthreads();
    // truncation box of the centre tables + box for (Psi(tau;phi) - Psi(tau;x)) W on [0,hs]; same for the derivative
    if (frames && tid < N * N) {
        const int a = tid / N, bb = tid % N;
        double sm = 0;
#pragma unroll
        for (int cc = 0; cc < D; cc++) sm = __dadd_ru(sm, imag(S.Ws[cc][bb]));
        double s0 = __dmul_ru(S.mj.rho_psi, sm), s1 = __dmul_ru(S.mj.rho_psi_d, sm);
#pragma unroll
        for (int cc = 0; cc < D; cc++) { const double wm = imag(S.Ws[cc][bb]); s0 = __fma_ru(PT.magXu[a][cc], wm, s0); s1 = __fma_ru(PT.magXv[a][cc], wm, s1); }
        S.fk()[0][a][bb] = iadd(S.fk()[0][a][bb], sym(s0));
        S.boxFd[a][bb] = s1;
    }
    __syncthreads();
}

// ---- set update (oracle/c2.hpp, same products and the same order of additions).  Called by all threads of the CTA.
//   front:  C+ = mid(J C),  r+ = J r + (y - x+) + (J C - C+) r0      with J = Jc + [0,1] dJ
//   frame:  T+ = (Jc + sum_j Xi_j r0_j + Xi_r)(Tc + sum_l Theta_l r0_l + R):  Tc+ = mid(Jc Tc), Theta_j+ = mid(Jc Theta_j + Xi_j Tc),
//           R+ = everything else;  Ps+ = J Ps.  Products with a point factor are fused interval*point accumulations (one rounding per
//           bound and term: the rounding noise of Jc Tc goes straight into R); the rest are (mid,t) dot products.
//   New values stay in registers until every thread has read the old state.  Returns the non-finite flag. ----
template <int N>
__device__ __noinline__ int set_update(FrontSmem<N>& S, PostTab<N>& PT, int cur, int nxt) {
    constexpr int D = 2 * N;
    const int tid = threadIdx.x ^ S.vswap;
    if (tid < D) S.x[nxt][tid] = __dadd_ru(S.x[cur][tid], mid_ru(S.y[tid]));          // x+ = x + mid(increment)
    // item loops are rolled and every item type appears once (code size); new values are staged in the dead table space and
    // copied back after the barrier.  Thread offsets spread the four item types over the CTA.
    constexpr int NJC = D * D, NTH = D * D * N, NR = D * N;
    constexpr int OFF_R = NTH % NT, OFF_PS = (OFF_R + NR) % NT, OFF_JC = (OFF_PS + NR) % NT;
#pragma unroll 1
    for (int t = tid; t < NTH; t += NT) {
        const int j = t / (D * N), i = (t / N) % D, k = t % N;
        ival a = pt(0.0);
#pragma unroll
        for (int l = 0; l < D; l++) ipfma(a, S.Jc[i][l], S.Th[j][l][k]);
#pragma unroll
        for (int l = 0; l < D; l++) ipfma(a, PT.Xi[j][i][l], S.Tc[l][k]);
        const double th = mid_ru(a);
        PT.u.e.Thn[j][i][k] = th;
        PT.u.e.E[j][i][k] = imul(isub(a, pt(th)), S.r0[j]);
    }
#pragma unroll 1
    for (int t = (tid + NT - OFF_R) % NT; t < NR; t += NT) {
        const int i = t / N, k = t % N;
        ival jt = pt(0.0);
#pragma unroll
        for (int l = 0; l < D; l++) ipfma(jt, S.Kc[i][l], S.Tc[l][k]);                 // increment (DPhi(x) - I) Tc
        const double tc0 = S.Tc[i][k], tc = __dadd_ru(tc0, mid_ru(jt));
        jt = iadd(isub(pt(tc0), pt(tc)), jt);                                            // Jc Tc - Tc+  (thin)
        acc4 ea = acc_zero();
#pragma unroll
        for (int l = 0; l < D; l++) acc_term(ea, PT.Jcm[i][l], S.Rm[l][k]);
#pragma unroll
        for (int l = 0; l < D; l++) acc_term(ea, PT.dJm[i][l], S.Tdm[l][k]);
        ival xr = pt(0.0);
#pragma unroll
        for (int l = 0; l < D; l++) ipfma(xr, PT.Xi[D][i][l], S.Tc[l][k]);
        PT.u.e.Tcn[i][k] = tc;
        PT.u.e.Rn[i][k] = iadd(iadd(jt, acc_finish(ea)), xr);
    }
#pragma unroll 1
    for (int t = (tid + NT - OFF_PS) % NT; t < NR; t += NT) {
        const int i = t / N, k = t % N;
        acc4 a = acc_zero();
#pragma unroll
        for (int l = 0; l < D; l++) acc_term(a, PT.Jm[i][l], S.Psm[l][k]);
        PT.u.e.Psn[i][k] = acc_finish(a);
    }
#pragma unroll 1
    for (int t = (tid + NT - OFF_JC) % NT; t < NJC; t += NT) {
        const int i = t / D, j = t % D;
        ival a = pt(0.0);
#pragma unroll
        for (int k = 0; k < D; k++) ipfma(a, S.J[i][k], S.C[cur][k][j]);
        const double cm = mid_ru(a);
        S.C[nxt][i][j] = cm;
        S.un.JCr[i][j] = iv2mt(isub(a, pt(cm)));
    }
    __syncthreads();
    const int nonfinite = S.flag & 1;
#pragma unroll 1
    for (int t = tid; t < NTH; t += NT) (&S.Th[0][0][0])[t] = (&PT.u.e.Thn[0][0][0])[t];
    if (tid >= 32 && tid < 32 + D) {
        const int i = tid - 32;
        acc4 a = acc_zero();
#pragma unroll
        for (int j = 0; j < D; j++) acc_term(a, S.un.JCr[i][j], S.r0m[j]);
#pragma unroll
        for (int j = 0; j < D; j++) acc_term(a, PT.Jm[i][j], iv2mt(S.r[cur][j]));
        S.r[nxt][i] = iadd(iadd(isub(pt(S.x[cur][i]), pt(S.x[nxt][i])), S.y[i]), acc_finish(a));   // (x - x+) + increment + ...
    }
#pragma unroll 1
    for (int t = tid; t < NR; t += NT) {
        const int i = t / N, k = t % N;
        ival rr = PT.u.e.Rn[i][k];
#pragma unroll
        for (int j = 0; j < D; j++) rr = iadd(rr, PT.u.e.E[j][i][k]);
        S.R[i][k] = rr; S.Tc[i][k] = PT.u.e.Tcn[i][k]; S.Ps[i][k] = PT.u.e.Psn[i][k];
    }
    __syncthreads();
    return nonfinite;
}

// ---- basis renormalisation of the frame, once per unit of time (oracle/c2.hpp:renorm_matrix and the block after the set update,
//      same operations in the same order).  The columns of T = Dphi_t A_u align with the fastest unstable direction (condition number
//      ~1e5 at L_-): the interval determinant of the hull of T then loses every digit once the relative width of the entries exceeds
//      1/cond, which with a validated neighbourhood of the connecting orbit (1e-13 .. 1e-12) is the last third of the front.  T is only
//      a BASIS of E^u: det(top(T)) = det(top(T H)) / det(H) for any invertible point matrix H.  H = inverse R factor of a floating-point
//      Gram-Schmidt of Tc (round to nearest: any H is valid, only determinism matters), applied rigorously:
//        Tc <- mid([Tc H]),  Theta_j <- mid([Theta_j H]),  R <- R H + residuals,  Err <- Err H,  G <- [H^-1] G,  sc <- sc det[H^-1].
//      Called by all threads of the CTA. ----
template <int N>
__device__ __noinline__ void renormalise(FrontSmem<N>& S) {
    constexpr int D = 2 * N;
    const int tid = threadIdx.x ^ S.vswap;
    static_assert(D * N <= 32 && N * N + 1 <= 32, "renormalise: one item per thread");
    if (tid == 0) {
        double Q[D][N], Rt[N][N], H[N][N];
        bool ok = true;
        for (int k = 0; k < N; k++) {
            double v[D];
            for (int i = 0; i < D; i++) v[i] = S.Tc[i][k];
            for (int m = 0; m < k; m++) {
                double dp = 0.0;
                for (int i = 0; i < D; i++) dp = __dadd_rn(dp, __dmul_rn(Q[i][m], v[i]));
                Rt[m][k] = dp;
                for (int i = 0; i < D; i++) v[i] = __dadd_rn(v[i], -__dmul_rn(dp, Q[i][m]));
            }
            double nn = 0.0;
            for (int i = 0; i < D; i++) nn = __dadd_rn(nn, __dmul_rn(v[i], v[i]));
            nn = __dsqrt_rn(nn);
            if (!(nn > 0.0) || !isfinite(nn)) { ok = false; nn = 1.0; }
            Rt[k][k] = nn;
            for (int i = 0; i < D; i++) Q[i][k] = __ddiv_rn(v[i], nn);
        }
        for (int k = 0; k < N; k++) {
            for (int i = 0; i < N; i++) H[i][k] = 0.0;
            H[k][k] = __ddiv_rn(1.0, Rt[k][k]);
            for (int i = k - 1; i >= 0; i--) {
                double a = 0.0;
                for (int m = i + 1; m <= k; m++) a = __dadd_rn(a, __dmul_rn(Rt[i][m], H[m][k]));
                H[i][k] = __ddiv_rn(-a, Rt[i][i]);
            }
        }
        for (int k = 0; k < N; k++) for (int i = 0; i <= k; i++) if (!isfinite(H[i][k])) ok = false;
        for (int k = 0; k < N; k++) if (!(H[k][k] > 0.0)) ok = false;
        if (ok) {
            for (int k = 0; k < N; k++) {               // interval inverse of the upper triangular point matrix H
                for (int i = 0; i < N; i++) { S.un.rn.Hm[i][k] = H[i][k]; S.un.rn.Hi[i][k] = pt(0.0); }
                S.un.rn.Hi[k][k] = mk(__ddiv_rd(1.0, H[k][k]), __ddiv_ru(1.0, H[k][k]));
                for (int i = k - 1; i >= 0; i--) {
                    ival a = pt(0.0);
                    for (int m = i + 1; m <= k; m++) ipfma(a, S.un.rn.Hi[m][k], H[i][m]);
                    const ival na = ineg(a);
                    const double hd = H[i][i];          // > 0
                    S.un.rn.Hi[i][k] = mk(dmin(__ddiv_rd(na.lo, hd), __ddiv_rd(na.hi, hd)), dmax(__ddiv_ru(na.lo, hd), __ddiv_ru(na.hi, hd)));
                }
            }
        }
        S.un.rn.ok = ok ? 1 : 0;
    }
    __syncthreads();
    if (!S.un.rn.ok) { __syncthreads(); return; }
    // one item per thread, new values in registers until every thread has read the old state
    double tc2 = 0.0, th2[D]; ival r2 = pt(0.0), e2 = pt(0.0), g2 = pt(0.0);
    const int i = tid / N, k = tid % N;
    if (tid < D * N) {
        ival t = pt(0.0);
        for (int l = 0; l <= k; l++) ipfma(t, pt(S.Tc[i][l]), S.un.rn.Hm[l][k]);
        tc2 = mid_ru(t);
        ival rr = isub(t, pt(tc2));
        for (int l = 0; l <= k; l++) ipfma(rr, S.R[i][l], S.un.rn.Hm[l][k]);
#pragma unroll
        for (int j = 0; j < D; j++) {
            ival u = pt(0.0);
            for (int l = 0; l <= k; l++) ipfma(u, pt(S.Th[j][i][l]), S.un.rn.Hm[l][k]);
            th2[j] = mid_ru(u);
            rr = iadd(rr, imul_nl(isub(u, pt(th2[j])), S.r0[j]));
        }
        r2 = rr;
    } else if (tid >= 32 && tid < 32 + N * N) {
        const int jj = (tid - 32) / N, kk = (tid - 32) % N;
        ival t = pt(0.0);
        for (int l = 0; l <= kk; l++) ipfma(t, S.Err[jj][l], S.un.rn.Hm[l][kk]);
        e2 = t;
        ival a = pt(0.0);
        for (int l = jj; l < N; l++) a = iadd(a, imul_nl(S.un.rn.Hi[jj][l], S.G[l][kk]));
        g2 = a;
    } else if (tid == 32 + N * N) {
        ival c = S.sc;
        for (int q = 0; q < N; q++) c = imul_nl(c, S.un.rn.Hi[q][q]);
        S.sc = c;
        S.renorms = S.renorms + 1;
    }
    __syncthreads();
    if (tid < D * N) {
        S.Tc[i][k] = tc2; S.R[i][k] = r2;
#pragma unroll
        for (int j = 0; j < D; j++) S.Th[j][i][k] = th2[j];
    } else if (tid >= 32 && tid < 32 + N * N) {
        const int jj = (tid - 32) / N, kk = (tid - 32) % N;
        S.Err[jj][kk] = e2; S.Errm[jj][kk] = iv2mt(e2); S.G[jj][kk] = g2;
    }
    __syncthreads();
}

// phi' = (v, F(u)) over the sub-interval [tl,th] of the step and over the whole set (last column of the first / last frame,
// topFrame.cpp:265-331), row i:  phi(tau) in phi_c(tau) + Psi(tau;xi)(phi0 - x),  |Psi(tau;xi) - Psi_c(tau)| <= g1 tau |phi0 - x|_1
template <int N>
__device__ __noinline__ void phi_prime(const FrontSmem<N>& S, const mt* Tf, int i, int p, int cur, double hs, double tl, double th, const Majorant& mj, ival& du, ival& dv) {
    using L = Lay<N>;
    constexpr int D = 2 * N;
    ival dev[D]; double dn = 0;
    for (int q = 0; q < D; q++) {
        ival a = pt(0.0);
        for (int j = 0; j < D; j++) a = iadd(a, imul_nl(pt(S.C[cur][q][j]), S.r0[j]));
        dev[q] = iadd(a, S.r[cur][q]); dn = __dadd_ru(dn, imag(dev[q]));
    }
    const double pb = __dadd_ru(mj.rho_psi, __dmul_ru(__dmul_ru(mj.g1, hs), dn));
    ival uu[N], gg[N];
    for (int m = 0; m < N; m++) {
        ival a = iadd(eval_rg(Tf + L::U(m) * KMAX, p, tl, th), sym(mj.rho_phi));
        for (int cc = 0; cc < D; cc++) a = iadd(a, imul_nl(iadd(eval_rg(Tf + L::PU(m, cc) * KMAX, p, tl, th), sym(pb)), dev[cc]));
        uu[m] = a; gg[m] = imul_nl(a, isub(pt(1.0), a));
    }
    {
        ival bq = iadd(eval_rg(Tf + L::V(i) * KMAX, p, tl, th), sym(mj.rho_phi));
        for (int cc = 0; cc < D; cc++) bq = iadd(bq, imul_nl(iadd(eval_rg(Tf + L::PV(i, cc) * KMAX, p, tl, th), sym(pb)), dev[cc]));
        du = bq;
    }
    const ival wi = isub(uu[i], pt(0.5));
    ival f = ineg(imul_nl(imul_nl(gg[i], wi), S.bpar[i]));
    if (i > 0)     f = isub(f, imul_nl(imul_nl(gg[i - 1], wi), S.cpar[i - 1]));
    if (i < N - 1) f = isub(f, imul_nl(imul_nl(gg[i + 1], wi), S.cpar[i]));
    dv = f;
}

// Enclosure of F'(tau) over [tl, th] for every frame entry: Horner on the derivative series k*F_k, plus the tail bound
// (the reference's rangeEnclosure of the vector field along the frame, utils.cpp:243-251).  Cold: only sub-intervals whose
// rough test fails reach it.
template <int N>
__device__ __noinline__ void deriv_entries(const ival* __restrict__ Fk0, const double* __restrict__ rhoFd, int p, double tl, double th, ival (&Dm)[N][N]) {
    constexpr int NN = N * N;
    for (int a = 0; a < N; a++)
        for (int bb = 0; bb < N; bb++) {
            const int c = a * N + bb;
            ival acc = imulk(Fk0[p * NN + c], p);
            for (int k = p - 1; k >= 1; k--) acc = horner_rg(acc, tl, th, imulk(Fk0[k * NN + c], k));
            Dm[a][bb] = iadd(acc, sym(rhoFd[c]));
        }
}

// ---- per-lane constants of the table loop: fixed for the whole front (they depend on the lane and on n only).  Packed into a few words
//      that travel to table_loop in registers; unpacking is one shift/mask per use. ----
struct LaneC {
    TaskR tA, tB, tC, tD;    // phi warp: tA = this lane's A or B task;  Psi warp: tA/tB and tC/tD = this lane's pairs of C tasks (rounds 0, 1)
    unsigned cs;             // cs0 | cs1 << 8 | cs2 << 16 | gslot << 24       scratch slots (combine stages); gslot 255 = none
    unsigned ap;             // ap0 | ap1 << 6 | ap2 << 12 | mscale << 18 | mcst << 19 | (mrole + 1) << 20 | ccomb << 28 (lane combines the sums of one Psi entry) | cpat << 29
    unsigned rvu;            // rowV | rowU << 16       table offsets (row * KMAX)
    unsigned rz;             // rowZ
    __device__ __forceinline__ int cs0() const { return (int)(cs & 0xffu); }
    __device__ __forceinline__ int cs1() const { return (int)((cs >> 8) & 0xffu); }
    __device__ __forceinline__ int cs2() const { return (int)((cs >> 16) & 0xffu); }
    __device__ __forceinline__ int gslot() const { return (int)(cs >> 24); }
    __device__ __forceinline__ int ap0() const { return (int)(ap & 0x3fu); }
    __device__ __forceinline__ int ap1() const { return (int)((ap >> 6) & 0x3fu); }
    __device__ __forceinline__ int ap2() const { return (int)((ap >> 12) & 0x3fu); }
    __device__ __forceinline__ int mscale() const { return (int)((ap >> 18) & 1u); }
    __device__ __forceinline__ int mcst() const { return (int)((ap >> 19) & 1u); }
    __device__ __forceinline__ int mrole() const { return (int)((ap >> 20) & 0xffu) - 1; }
    __device__ __forceinline__ int ccomb() const { return (int)((ap >> 28) & 1u); }
    __device__ __forceinline__ int cpat() const { return (int)(ap >> 29); }          // singles: 0 = first component (no w = 1 product), 1 = middle, 2 = last (no w = 2 product)
    __device__ __forceinline__ int rowV() const { return (int)(rvu & 0xffffu); }
    __device__ __forceinline__ int rowU() const { return (int)(rvu >> 16); }
    __device__ __forceinline__ int rowZ() const { return (int)rz; }
};
template <int N> struct TLC {                      // scratch layout of the table loop
    using L = Lay<N>;
    static constexpr int ZSLOT = L::NB + L::NC;      // always-zero slot (missing chain neighbours)
    static constexpr int GS = ZSLOT + 1;             // scratch slots of g_i (N) and q_e (N-1) of the newest order
    static constexpr int MB = L::NA + L::NB;         // phi warp lanes: A tasks 0..NA-1 | B tasks NA..MB-1 | M/H lanes MB..MB+3N-2 | hosted C tasks HOST0..
    static_assert(MB + 3 * N - 1 <= 32, "phi warp: A, B and M/H lanes side by side");
    // Distribution of the 2N(3N-2) variational products (C tasks).  The products of one column cc of Psi only involve that column, so
    // whole columns can live in either warp without any exchange but the M-series.  n <= 3 ("singles"): every lane owns ONE product --
    // the Psi warp the columns HC..2N-1, the idle lanes of the phi warp the columns 0..HC-1 -- so both warps run the single-task pass
    // (8 DFMA per term; a pair pass costs 16 DFMA instructions per term however few lanes are busy, and at full occupancy the passes are
    // bound by the FP64 pipes).  n = 4: pairs of products sharing the M-series, two rounds, all in the Psi warp.
    static constexpr bool SINGLES = CCP_SINGLES && N <= 3;
    static constexpr int PPC = 3 * N - 2;            // products per column
    static constexpr int HOST0 = MB + 3 * N - 1;     // first hosted lane of the phi warp
    static constexpr int HC = !SINGLES ? 0 : (2 * N - 32 / PPC > 0 ? 2 * N - 32 / PPC : 0);      // hosted columns: those the Psi warp cannot hold (n = 3: 2 of 6)
    static_assert(!SINGLES || ((2 * N - HC) * PPC <= 31 && HOST0 + HC * PPC <= 32), "singles: one product per lane; lane 31 of the Psi warp stays idle (exact zero for entry_sum)");
    static_assert(GS + 2 * N - 1 < 255, "scratch slots fit 8 bits");
};
template <int N>
__device__ inline LaneC make_lanec(int tid) {
    using L = Lay<N>;
    constexpr int D = 2 * N, ZSLOT = TLC<N>::ZSLOT, GS = TLC<N>::GS, MB = TLC<N>::MB, PZ = 2 * N - 1;
    constexpr int CROUNDS = PhaseCheck<N>::CROUNDS;
    const int wq = tid >> 5, lane = tid & 31;
    LaneC c;
    constexpr bool SINGLES = TLC<N>::SINGLES;
    constexpr int PPC = TLC<N>::PPC, HOST0 = TLC<N>::HOST0, HC = TLC<N>::HC;
    // index of the product (i, cc, w) in the enumeration of decode_task (w = 0: Md_i Pu_i, 1: Mo_{i-1} Pu_{i-1}, 2: Mo_i Pu_{i+1})
    auto c_index = [](int i, int cc, int w) -> int {
        const int nw = (i == 0 || i == N - 1) ? 2 : 3;
        const int widx = (i == 0) ? (w == 0 ? 0 : 1) : w;                     // i = 0 has w in {0, 2}; the others start at w = 0, 1
        return (i == 0 ? 0 : (3 * i - 1) * D) + cc * nw + widx;
    };
    // k-th product of a column, i-major: (i, w) pairs in the order (0,0) (0,2) (1,0) (1,1) (1,2) ... (N-1,0) (N-1,1)
    auto column_product = [&](int cc, int k) -> int {
        int cnt = 0, idx = 1 << 20;
        for (int i = 0; i < N; i++) for (int w = 0; w < 3; w++) {
            if ((w == 1 && i == 0) || (w == 2 && i == N - 1)) continue;
            if (cnt == k) idx = c_index(i, cc, w);
            cnt++;
        }
        return idx;
    };
    int pi0 = 1 << 20, pi1 = 1 << 20, pi2 = 1 << 20, pi3 = 1 << 20;
    if (SINGLES) {
        if (wq == 1 && lane < (D - HC) * PPC) pi0 = column_product(HC + lane / PPC, lane % PPC);                  // Psi warp: columns HC..D-1
        if (wq == 0 && lane >= HOST0 && lane < HOST0 + HC * PPC) pi0 = column_product((lane - HOST0) / PPC, (lane - HOST0) % PPC);   // phi warp: hosted columns
    } else {
        pair_tasks<N>(wq == 1 ? psi_lane_to_pair<N>(lane) : -1, pi0, pi1);
        pair_tasks<N>((wq == 1 && CROUNDS > 1) ? 32 + lane : -1, pi2, pi3);
    }
    const bool hosted = SINGLES && wq == 0 && lane >= HOST0;
    c.tA = make_taskr((wq == 0 && !hosted) ? (lane < L::NA ? decode_task<N>(0, lane) : decode_task<N>(1, lane - L::NA)) : decode_task<N>(2, pi0));
    c.tB = make_taskr(decode_task<N>(2, wq == 0 ? (1 << 20) : pi1));
    c.tC = make_taskr(decode_task<N>(2, pi2)); c.tD = make_taskr(decode_task<N>(2, pi3));
    // operands of the "combine" stages.  Every stage is ONE branch-free instruction stream (missing neighbours: zero parameter, clamped slot).
    //   phi warp, lanes MB..    : one (mid,t) dot product of three parameter * series slots (parameter tables S.pm, g / q of the newest order from
    //                             the A lanes through scratch):  Md_i[k] = [k==0] b_i/2 - (3 b_i g_i + c g_{i-1} + c g_{i+1}),
    //                             Mo_e[k] = 2 c_e q_e,  H_i[k] = -(b_i g_i + c g_{i-1} + c g_{i+1})   (v' = H w)
    //   Psi warp, lanes < N*D   : Pv[k] = (X0 + X1 + X2)/k from the C sums (scratch); Pu[k] = Pv[k-1]/k
    //   The B lanes of the phi warp turn their own Cauchy sum (H w)[k] into v[k+1], u[k+2], z[k+2].
    int cs0 = ZSLOT, cs1 = ZSLOT, cs2 = ZSLOT, rowV = 0, rowU = 0, rowZ = 0;
    int ap0 = PZ, ap1 = PZ, ap2 = PZ, mscale = 1, gslot = 255, mcst = 0;
    const int mrole = (wq == 0 && lane >= MB && lane < MB + 3 * N - 1) ? lane - MB : -1;      // < N: Md_i, < 2N-1: Mo_e, else H_i
    if (wq == 0) {
        if (lane >= L::NA && lane < MB) {                    // B lane: v, u, z rows of component i
            const int i = lane - L::NA;
            rowV = L::V(i) * KMAX; rowU = L::U(i) * KMAX; rowZ = L::Z(i) * KMAX;
        } else if (mrole >= 0 && (mrole < N || mrole >= 2 * N - 1)) {   // Md_i, H_i
            const bool isH = mrole >= 2 * N - 1;
            const int i = isH ? mrole - (2 * N - 1) : mrole;
            cs0 = GS + i; cs1 = GS + (i > 0 ? i - 1 : 0); cs2 = GS + (i < N - 1 ? i + 1 : N - 1);       // clamped: missing neighbours carry zero parameters
            ap0 = (isH ? 0 : N) + i; ap1 = 3 * N + i; ap2 = 4 * N + i;                                    // B or B3 | Cm | Cp in S.pm
            mscale = 1; mcst = isH ? 0 : 1; rowV = (isH ? L::G(i) : L::MD(i)) * KMAX;                    // mscale: negate;  mcst: add b_i/2 at order 0
        } else if (mrole >= N) {                             // Mo_e = 2 c_e q_e
            const int e = mrole - N;
            cs0 = cs1 = cs2 = GS + N + e; ap0 = 7 * N + e; ap1 = ap2 = 3 * N;                             // C2 | zero (Cm[0]) | zero
            mscale = 0; rowV = L::MO(e) * KMAX;
        }
        if (lane < L::NA) gslot = GS + lane;                 // A lanes: g (task i) and q (task N+e)
    }
    // lanes that combine the sums of one entry (i, cc) of Psi: Psi warp: the entries of its columns; phi warp (singles): the hosted columns
    int ccomb = 0, cpat = 0;
    {
        int e = -1, ci = 0, ccc = 0;
        if (SINGLES) {
            // singles: the lane that owns the FIRST product of an entry (i, cc) combines the entry; its partners are the next one or two lanes
            // (products of a column in the order (0,0) (0,2) | (1,0) (1,1) (1,2) | ... | (N-1,0) (N-1,1)), whose sums arrive through shuffles
            int pl = -1, c0l = 0;
            if (wq == 1 && lane < (D - HC) * PPC) { pl = lane % PPC; c0l = HC + lane / PPC; }
            if (wq == 0 && lane >= HOST0 && lane < HOST0 + HC * PPC) { pl = (lane - HOST0) % PPC; c0l = (lane - HOST0) / PPC; }
            if (pl == 0) { e = 0; ci = 0; ccc = c0l; }
            else if (pl >= 2 && (pl - 2) % 3 == 0) { e = 0; ci = 1 + (pl - 2) / 3; ccc = c0l; }
            cpat = ci == 0 ? 0 : (ci == N - 1 ? 2 : 1);
            // source lanes of the two partner sums (entry_sum): a product a boundary component does not have is read from a lane whose sum is
            // an exact zero -- lane 31 of the Psi warp (never owns a product), lane 0 of the phi warp (an A lane) -- instead of being selected.
            // The B lanes read two zeros: their "entry sum" is their own sum (x + 0 is exact and keeps the sign of a zero bound: the upper
            // bound of a finished sum is never -0), so B lanes and entry leaders share the epilogue without a select.  The lanes that use
            // scratch slots (M/H lanes) are neither: cs0, cs1 carry the two source lanes.
            const int zl = wq == 1 ? 31 : 0;
            if (e >= 0) { cs0 = cpat == 0 ? zl : lane + 1; cs1 = cpat == 0 ? lane + 1 : (cpat == 1 ? lane + 2 : zl); }
            else if (wq == 0 && lane >= L::NA && lane < MB) { cs0 = zl; cs1 = zl; }
        } else if (wq == 1 && lane < N * D) { e = lane; ci = e / D; ccc = e % D; }
        if (e >= 0) {
            const int i = ci, cc = ccc;
            if (!SINGLES) {
                int c0 = L::NB + (i == 0 ? 0 : (3 * i - 1) * D) + cc * ((i == 0 || i == N - 1) ? 2 : 3);
                cs0 = c0++;
                if (i > 0)     cs1 = c0++;
                if (i < N - 1) cs2 = c0;
            }
            rowV = L::PV(i, cc) * KMAX; rowU = L::PU(i, cc) * KMAX;
            ccomb = 1;
        }
    }
    c.cs = opaqueu((unsigned)cs0 | ((unsigned)cs1 << 8) | ((unsigned)cs2 << 16) | ((unsigned)gslot << 24));
    c.ap = opaqueu((unsigned)ap0 | ((unsigned)ap1 << 6) | ((unsigned)ap2 << 12) | ((unsigned)mscale << 18) | ((unsigned)mcst << 19) | ((unsigned)(mrole + 1) << 20) | ((unsigned)ccomb << 28) | ((unsigned)(ccomb ? cpat : 0) << 29));
    c.rvu = opaqueu((unsigned)rowV | ((unsigned)rowU << 16));
    c.rz = opaqueu((unsigned)rowZ);
    return c;
}

__device__ __forceinline__ ival shfl_ival(const ival& a, int src) {
    return mk(__shfl_sync(0xffffffffu, a.lo, src), __shfl_sync(0xffffffffu, a.hi, src));
}

// The lane constants are the same for every front of one n: a table in global memory, filled once per device (lanec_table_kernel, launched
// by ccp_init) and read through the read-only path at the start of every table loop -- 64 bytes per lane instead of 16 registers that
// would be live across the whole step.
__device__ uint4 g_lanec[CCP_MAX_N + 1][NT][4];
template <int N> __global__ void lanec_table_kernel() {
    const LaneC c = make_lanec<N>(threadIdx.x);
    uint4* o = g_lanec[N][threadIdx.x];
    o[0] = make_uint4(c.tA.ab, c.tA.zz, c.tA.ok, c.tB.ab); o[1] = make_uint4(c.tB.zz, c.tB.ok, c.tC.ab, c.tC.zz);
    o[2] = make_uint4(c.tC.ok, c.tD.ab, c.tD.zz, c.tD.ok); o[3] = make_uint4(c.cs, c.ap, c.rvu, c.rz);
}

// ---- Taylor tables of the centre trajectory, two orders per stage (called by all threads of the CTA once per step; its own function so
//      that the register allocation of the hot loop is not shared with the long-lived state of the step loop).
//      u'' = F(u): the coefficients of order k+2 depend on those of order <= k only, so the orders J, J+1 (J even) are independent of each
//      other and are produced together; the stages form the dependent chain (10 of them for p = 20 instead of 20 orders).  Stage (J, J+1):
//        phi warp:  pass: A lanes full sums g, q [J], [J+1];  B lanes the part of (H w)[J], (H w)[J+1] that does not need H[J], H[J+1]
//                   -> M/H lanes: Md, Mo, H of both orders  -> (bar: the M-series travel to the Psi warp)
//                   -> B lanes: the terms with H[J], H[J+1], then v[J+1], v[J+2], u[J+2], u[J+3], z
//        Psi warp:  pass: the part of (M Psi^u)[J], [J+1] that does not need M[J], M[J+1]  -> (bar)  -> the terms with M[J], M[J+1]
//                   -> Pv[J+1], Pv[J+2], Pu[J+2], Pu[J+3]
//      Every sum keeps the summation order of oracle/c1.hpp:cauchy_order (interior terms ascending, then the end terms j = 0 and
//      j = order): only the schedule differs from an order-by-order loop, the values are bit-identical. ----
template <int N>
__device__ __noinline__ void table_loop(FrontSmem<N>& S) {
    LaneC lc;
    {
        const uint4* q = g_lanec[N][threadIdx.x ^ S.vswap];
        const uint4 w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3);
        lc.tA.ab = w0.x; lc.tA.zz = w0.y; lc.tA.ok = w0.z; lc.tB.ab = w0.w; lc.tB.zz = w1.x; lc.tB.ok = w1.y;
        lc.tC.ab = w1.z; lc.tC.zz = w1.w; lc.tC.ok = w2.x; lc.tD.ab = w2.y; lc.tD.zz = w2.z; lc.tD.ok = w2.w;
        lc.cs = w3.x; lc.ap = w3.y; lc.rvu = w3.z; lc.rz = w3.w;
    }
    const int p = S.cfg.p;
    constexpr int GS = TLC<N>::GS;
    constexpr int CROUNDS = PhaseCheck<N>::CROUNDS;
    constexpr bool SINGLES = TLC<N>::SINGLES;
    constexpr int HC = TLC<N>::HC;
    const int tid = threadIdx.x ^ S.vswap, wq = tid >> 5;
    mt (*T)[KMAX] = S.T;
    const mt* const Tf = &T[0][0];
#define cs0 lq.cs0()
#define cs1 lq.cs1()
#define cs2 lq.cs2()
#define ap0 lq.ap0()
#define ap1 lq.ap1()
#define ap2 lq.ap2()
#define rowV lq.rowV()
#define rowU lq.rowU()
#define rowZ lq.rowZ()
#define gslot lq.gslot()
#define mscale lq.mscale()
#define mcst lq.mcst()
#define mrole lq.mrole()
#define ccomb lq.ccomb()
#define cpat lq.cpat()
#undef PROF
#undef TRC
#define PROF(i) do {} while (0)
#define TRC(i) do {} while (0)
        ival* const scr0 = S.un.scratch;                     // sums / g, q of order J
        ival* const scr1 = &S.Jc[0][0];                      // ... of order J+1 (Jc, Kc are dead during the table loop)
        mt* const dummy = reinterpret_cast<mt*>(scr1) + GS + 2 * N - 1;        // target of the stores that fall beyond order p
        const mt* const inv = reinterpret_cast<const mt*>(scr1) + FrontSmem<N>::INV_OFF;   // (RU(1/k), RU(RU(1/k)(1 + 2^-50))), refilled every step (Jc, Kc, J are dead here)
        static_assert(FrontSmem<N>::INV_OFF >= GS + 2 * N, "second scratch set + dummy slot, then the reciprocal table");
        auto stage = [&](auto fast_tag, const int J) {
            constexpr int MODE = decltype(fast_tag)::value;     // 1: J >= 2 and two orders, 2: J == 0 and two orders, 0: decided at run time
            const bool two = MODE != 0 || (J + 1 < p);            // uniform: the stage has a second order
            const bool three = J + 3 <= p;                        // u, z, Pu of order J+3 exist
            // the packed lane constants are laundered once per stage: everything derived from them (table addresses, scratch slots) is
            // recomputed here with a few ALU operations instead of being hoisted out of the stage loop into registers that spill
            LaneC lq = lc;
            lq.tA.ab = opaqueu(lq.tA.ab); lq.tA.zz = opaqueu(lq.tA.zz); lq.tA.ok = opaqueu(lq.tA.ok);
            lq.tB.ab = opaqueu(lq.tB.ab); lq.tB.zz = opaqueu(lq.tB.zz); lq.tB.ok = opaqueu(lq.tB.ok);
            if (CROUNDS > 1) {
                lq.tC.ab = opaqueu(lq.tC.ab); lq.tC.zz = opaqueu(lq.tC.zz); lq.tC.ok = opaqueu(lq.tC.ok);
                lq.tD.ab = opaqueu(lq.tD.ab); lq.tD.zz = opaqueu(lq.tD.zz); lq.tD.ok = opaqueu(lq.tD.ok);
            }
            lq.cs = opaqueu(lq.cs); lq.ap = opaqueu(lq.ap); lq.rvu = opaqueu(lq.rvu); lq.rz = opaqueu(lq.rz);
            const TaskR tA = lq.tA, tB = lq.tB, tC = lq.tC, tD = lq.tD;
            (void)tB; (void)tC; (void)tD;                         // pairs only
            // the lanes that own an entry (i, cc) of Psi: scratch sums of order kc -> Pv[kc+1];  Pu[kc+2] = Pv[kc+1]/(kc+2)   (Pu[1] = Pv[0] comes with the order-0 data)
            auto combine_c = [&]() {
                if (ccomb) {
                    const mt i1 = inv[J + 1], i2 = inv[J + 2], i3 = inv[J + 3];
                    const ival x0 = scr0[cs0], x1 = scr0[cs1], x2 = scr0[cs2], y0 = scr1[cs0], y1 = scr1[cs1], y2 = scr1[cs2];
                    const mt pv1 = mt_scale(iv2mt(iadd(iadd(x0, x1), x2)), i1.m, i1.t), pv2 = mt_scale(iv2mt(iadd(iadd(y0, y1), y2)), i2.m, i2.t);
                    const mt pu2 = mt_scale(pv1, i2.m, i2.t), pu3 = mt_scale(pv2, i3.m, i3.t);
                    T[0][rowV + J + 1] = pv1;
                    (two ? &T[0][rowV + J + 2] : dummy)[0] = pv2;
                    (two ? &T[0][rowU + J + 2] : dummy)[0] = pu2;
                    (three ? &T[0][rowU + J + 3] : dummy)[0] = pu3;
                }
            };
            // singles: the sums of the one or two partner lanes arrive through shuffles (every lane of the warp executes them); the leader of an
            // entry forms (w0 + w1) + w2 with an exact zero for the product a boundary component does not have -- the additions of combine_c
            auto entry_sum = [&](const ival& f) -> ival {
                const int sb = cs0 & 31, sc = cs1 & 31;               // make_lanec: partner lanes, or a lane holding an exact zero (M/H and A lanes: any lane, their result is not used)
                const ival b = shfl_ival(f, sb), c = shfl_ival(f, sc);
                return iadd(iadd(f, b), c);
            };
            if (wq == 0) {
                PROF(12);
                TRC(100 + J);
                const int kind = tA.kind();
                acc4 s0, s1;
                s0 = acc_zero(); s1 = acc_zero();                 // defined on every path (lanes without a task keep the zeros): accumulators that are only
                if (kind >= 0) pass_early<MODE>(Tf, tA, J, s0, s1);      // defined inside a divergent branch get spilled across the barrier below
                if (kind == 0) {
                    acc4 t0 = s0, t1 = s1;                        // A lanes: g, q of both orders -> scratch (they only travel to the M/H lanes)
                    pass_late<MODE>(Tf, tA, J, two, t0, t1);
                    const mt g0 = iv2mt(acc_finish(t0)), g1 = iv2mt(acc_finish(t1));
                    reinterpret_cast<mt*>(scr0)[gslot] = g0;
                    (two ? reinterpret_cast<mt*>(scr1) + gslot : dummy)[0] = g1;
                }
                __syncwarp();
                PROF(8);
                TRC(300 + J);
                ival f0 = pt(0.0), f1 = pt(0.0);
                if (mrole >= 0) {                                 // M/H lanes: one (mid,t) dot product per order: parameter * g (or q), three slots
                    const mt* const pmt = &S.pm.B[0];             // B | B3 | B6 | Cm | Cp | C2m | C2p | C2, N entries each
                    const mt p0 = pmt[ap0], p1 = pmt[ap1], p2 = pmt[ap2];
                    const mt* const gq0 = reinterpret_cast<const mt*>(scr0); const mt* const gq1 = reinterpret_cast<const mt*>(scr1);
                    acc4 c0 = acc_zero(), c1 = acc_zero();
                    acc_term(c0, p0, gq0[cs0]); acc_term(c1, p0, gq1[cs0]);
                    acc_term(c0, p1, gq0[cs1]); acc_term(c1, p1, gq1[cs1]);
                    acc_term(c0, p2, gq0[cs2]); acc_term(c1, p2, gq1[cs2]);
                    ival d0 = acc_finish(c0), d1 = acc_finish(c1);
                    if (mscale) { d0 = ineg(d0); d1 = ineg(d1); }
                    if ((MODE == 2 || (MODE == 0 && J == 0)) && mcst) d0 = iadd(S.pm.halfB[mrole], d0);      // Md_i[0] = b_i/2 - (...)
                    T[0][rowV + J] = iv2mt(d0);
                    (two ? &T[0][rowV + J + 1] : dummy)[0] = iv2mt(d1);
                }
                __syncwarp();
                TRC(400 + J);
                bar_pair();                                       // Md, Mo of orders J, J+1 are in the tables
                TRC(500 + J);
                if (kind == 1 || (SINGLES && HC > 0 && kind == 2)) {      // B lanes and hosted C lanes: the same late terms in one instruction stream
                    pass_late<MODE>(Tf, tA, J, two, s0, s1);
                    f0 = acc_finish(s0); f1 = acc_finish(s1);
                }
                if (SINGLES && HC > 0) {
                    // B lanes and the leaders of the hosted entries of Psi share ONE epilogue: sum -> (mid,t) -> /(kc+1) -> /(kc+2) -> rows V, U (B: and Z)
                    const ival e0 = entry_sum(f0), e1 = entry_sum(f1);
                    if (kind == 1 || ccomb) {
                        const ival g0 = e0, g1 = e1;                  // B lanes: e = f + 0 + 0 = f
                        const mt i1 = inv[J + 1], i2 = inv[J + 2], i3 = inv[J + 3];
                        const mt v1 = mt_scale(iv2mt(g0), i1.m, i1.t), v2 = mt_scale(iv2mt(g1), i2.m, i2.t);
                        const mt u2 = mt_scale(v1, i2.m, i2.t), u3 = mt_scale(v2, i3.m, i3.t);
                        mt z2, z3; z2.m = -u2.m; z2.t = u2.t; z3.m = -u3.m; z3.t = u3.t;
                        T[0][rowV + J + 1] = v1;
                        (two ? &T[0][rowV + J + 2] : dummy)[0] = v2;
                        (two ? &T[0][rowU + J + 2] : dummy)[0] = u2;
                        (three ? &T[0][rowU + J + 3] : dummy)[0] = u3;
                        ((two && kind == 1) ? &T[0][rowZ + J + 2] : dummy)[0] = z2;
                        ((three && kind == 1) ? &T[0][rowZ + J + 3] : dummy)[0] = z3;
                    }
                } else if (kind == 1) {                           // B lanes: (H w)[kc] is v'[kc]:  v[kc+1] = sum/(kc+1), u[kc+2] = v[kc+1]/(kc+2), z = -u
                    const mt i1 = inv[J + 1], i2 = inv[J + 2], i3 = inv[J + 3];
                    const mt v1 = mt_scale(iv2mt(f0), i1.m, i1.t), v2 = mt_scale(iv2mt(f1), i2.m, i2.t);
                    const mt u2 = mt_scale(v1, i2.m, i2.t), u3 = mt_scale(v2, i3.m, i3.t);
                    mt z2, z3; z2.m = -u2.m; z2.t = u2.t; z3.m = -u3.m; z3.t = u3.t;
                    T[0][rowV + J + 1] = v1;
                    (two ? &T[0][rowV + J + 2] : dummy)[0] = v2;
                    (two ? &T[0][rowU + J + 2] : dummy)[0] = u2;  (two ? &T[0][rowZ + J + 2] : dummy)[0] = z2;
                    (three ? &T[0][rowU + J + 3] : dummy)[0] = u3; (three ? &T[0][rowZ + J + 3] : dummy)[0] = z3;
                }
                __syncwarp();
                PROF(9);
            } else {
                TRC(700 + J);
                // variational tables: products M * Psi^u of orders J, J+1 (pairs of tasks sharing the M-series)
                if constexpr (SINGLES) {
                    // one product per lane: the single-task pass of the phi warp
                    const bool r0 = tA.kind() == 2;
                    acc4 s0 = acc_zero(), s1 = acc_zero();        // defined on every path (see the phi warp)
                    if (r0) pass_early<MODE>(Tf, tA, J, s0, s1);
                    TRC(400 + J);
                    bar_pair();
                    TRC(500 + J);
                    ival f0 = pt(0.0), f1 = pt(0.0);
                    if (r0) {
                        pass_late<MODE>(Tf, tA, J, two, s0, s1);
                        f0 = acc_finish(s0); f1 = acc_finish(s1);
                    }
                    TRC(800 + J);
                    const ival e0 = entry_sum(f0), e1 = entry_sum(f1);
                    if (ccomb) {                                  // Pv[kc+1] = sum/(kc+1);  Pu[kc+2] = Pv[kc+1]/(kc+2)   (Pu[1] = Pv[0] comes with the order-0 data)
                        const mt i1 = inv[J + 1], i2 = inv[J + 2], i3 = inv[J + 3];
                        const mt pv1 = mt_scale(iv2mt(e0), i1.m, i1.t), pv2 = mt_scale(iv2mt(e1), i2.m, i2.t);
                        const mt pu2 = mt_scale(pv1, i2.m, i2.t), pu3 = mt_scale(pv2, i3.m, i3.t);
                        T[0][rowV + J + 1] = pv1;
                        (two ? &T[0][rowV + J + 2] : dummy)[0] = pv2;
                        (two ? &T[0][rowU + J + 2] : dummy)[0] = pu2;
                        (three ? &T[0][rowU + J + 3] : dummy)[0] = pu3;
                    }
                    __syncwarp();
                } else {
                acc4 a0, a1, b0, b1, c0, c1, d0, d1;
                const bool r0 = tA.kind() == 2, r1 = CROUNDS > 1 && tC.kind() == 2;
                // every lane runs the passes (lanes without a task: on row 0, results dropped into the dummy slot) -- no divergence, and the
                // accumulators that live across the barrier are defined unconditionally
                ival* const dmy = reinterpret_cast<ival*>(dummy);
                a0 = acc_zero(); a1 = acc_zero(); b0 = acc_zero(); b1 = acc_zero(); c0 = acc_zero(); c1 = acc_zero(); d0 = acc_zero(); d1 = acc_zero();
                if (r0) pair_early<MODE>(Tf, tA, tB, J, a0, a1, b0, b1);
                if (r1) pair_early<MODE>(Tf, tC, tD, J, c0, c1, d0, d1);
                TRC(400 + J);
                bar_pair();
                TRC(500 + J);
                if (r0) {
                    pair_late<MODE>(Tf, tA, tB, J, two, a0, a1, b0, b1);
                    const ival fa0 = acc_finish(a0), fb0 = acc_finish(b0), fa1 = acc_finish(a1), fb1 = acc_finish(b1);
                    (r0 ? scr0 + tA.out() : dmy)[0] = fa0; (r0 ? scr0 + tB.out() : dmy)[0] = fb0;
                    (r0 ? scr1 + tA.out() : dmy)[0] = fa1; (r0 ? scr1 + tB.out() : dmy)[0] = fb1;      // scr1 sums of a missing order are never used
                }
                if (r1) {
                    pair_late<MODE>(Tf, tC, tD, J, two, c0, c1, d0, d1);
                    const ival fc0 = acc_finish(c0), fd0 = acc_finish(d0), fc1 = acc_finish(c1), fd1 = acc_finish(d1);
                    (r1 ? scr0 + tC.out() : dmy)[0] = fc0; (r1 ? scr0 + tD.out() : dmy)[0] = fd0;
                    (r1 ? scr1 + tC.out() : dmy)[0] = fc1; (r1 ? scr1 + tD.out() : dmy)[0] = fd1;
                }
                __syncwarp();
                TRC(800 + J);
                combine_c();
                __syncwarp();
                }
            }
        };

        for (int J = 0; J < p; J += 2) {
            if (J + 1 < p) { if (J >= 2) stage(cuda::std::integral_constant<int, 1>{}, J); else stage(cuda::std::integral_constant<int, CCP_STAGE0>{}, 0); }
            else stage(cuda::std::integral_constant<int, 0>{}, J);
        }
#undef cs0
#undef cs1
#undef cs2
#undef ap0
#undef ap1
#undef ap2
#undef rowV
#undef rowU
#undef rowZ
#undef gslot
#undef mscale
#undef mcst
#undef mrole
#undef ccomb
#undef cpat
#undef PROF
#undef TRC
}
#ifdef CCP_TRACE
#define TRC(id) do { if (blockIdx.x == 0 && s == 500 && (threadIdx.x & 31) == 0 && trc_n < 255) { g_trace[threadIdx.x >> 5][2 * trc_n] = clock64(); g_trace[threadIdx.x >> 5][2 * trc_n + 1] = (id); trc_n++; g_trace_n[threadIdx.x >> 5] = trc_n; } } while (0)
#else
#define TRC(id) do {} while (0)
#endif
#ifdef CCP_PROFILE
#define PROF(i) do { long long prof_n = clock64(); if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_prof[i], (unsigned long long)(prof_n - prof_t)); prof_t = prof_n; } while (0)
#else
#define PROF(i) do {} while (0)
#endif

// ------------------------------------------------------------------------------------------------
// One integrator step of one front (all threads of the CTA).  Its own function, with the front's configuration and tallies in shared
// memory, so that no long-lived state of the front occupies registers across the phases of a step.  Returns the status.
// ------------------------------------------------------------------------------------------------
template <int N>
__device__ __noinline__ int step_body(FrontSmem<N>& S, const int s, const unsigned align_sm, const int align_phase) {
    using L = Lay<N>;
    constexpr int D = 2 * N;
    const int tid = threadIdx.x ^ S.vswap, wq = tid >> 5, lane = tid & 31;
    mt (*T)[KMAX] = S.T;
    const mt* const Tf = &T[0][0];
    const Majorant& mj = S.mj;
    // PT overlays the Taylor tables once they are dead: second-variation data, EV[arg][row][col] frame enclosures, DT[.] determinants
    PostTab<N>& PT = *reinterpret_cast<PostTab<N>*>(&T[0][0]);
    ival* const EV = PT.u.EV;
    ival* const DT = PT.DT;
    const int p = S.cfg.p, g = S.cfg.g, Stot = S.cfg.Stot;
    const double L_minus = S.cfg.L_minus, h = S.cfg.h;
    ccp_front_result* const gres = S.cfg.gres;
    const ccp_front_desc* const gdesc = S.cfg.gdesc;
    ccp_series_rec* const series = S.cfg.series;
    const int series_cap = S.cfg.series_cap;
    const int cur = s & 1, nxt = cur ^ 1;
    int status = CCP_OK;
    {
        PROF_T0();
        TRC_DECL();
        TRC(1);
        if (align_sm != ~0u && align_phase == 0) { if (tid == 0) sm_align_wait(align_sm); __syncthreads(); }    // once per step: every 2nd step or twice per step are both slower (profiles/r2_v22_align_experiments.txt)
        const bool last = (s == Stot - 1);
        const double hs = last ? __dadd_ru(L_minus, -S.st.tprev.hi) : h;
        // grid points of this step: gp[i] = inf(i*h/grid), gp[grid] = h  (same for every full step)
        if ((s == 0 || last) && tid >= 32 && tid <= 32 + g) { const int i = tid - 32; S.gp[i] = (i < g) ? __ddiv_rd(__dmul_rd((double)i, hs), (double)g) : hs; }
        if (tid >= 32 && tid < 32 + KMAX + 2) { mt v; v.m = c_invhi[tid - 32]; v.t = c_invs[tid - 32]; reinterpret_cast<mt*>(&S.Jc[0][0])[FrontSmem<N>::INV_OFF + tid - 32] = v; }      // reciprocal table of the table loop (Jc, Kc, J are dead until the step image)
        if (tid == NT - 1) { S.un.scratch[TLC<N>::ZSLOT] = pt(0.0); (&S.Jc[0][0])[TLC<N>::ZSLOT] = pt(0.0); }     // zero slots of both scratch sets (the union was reused by the set update, Jc is rewritten every step)
        // ---- hull of the phi-set ----
        if (tid < D) {
            ival a = pt(S.x[cur][tid]);
            for (int j = 0; j < D; j++) ipfma(a, S.r0[j], S.C[cur][tid][j]);          // fused: one rounding per bound and term
            a = iadd(a, S.r[cur][tid]);
            S.X[tid] = a;
            if (!ifinite(a)) atomicOr(&S.flag, 1);
            else if (!isubset(a, tid < N ? mj.Gs_u : mj.Gs_v)) atomicOr(&S.flag, 2);
        }
        __syncthreads();
        {
            const int fl = S.flag;
            if (fl & 1) return CCP_ST_NONFINITE;
            if (fl & 2) return CCP_ST_ENCLOSURE;
        }
        // ---- order 0 of all tables (centre x); Psi = identity ----
        if (tid < N) {
            const int i = tid;
            const ival u0 = pt(S.x[cur][i]), v0 = pt(S.x[cur][N + i]);
            T[L::U(i)][0] = iv2mt(u0); T[L::V(i)][0] = iv2mt(v0);
            T[L::G(i)][KMAX - 1] = iv2mt(isub(u0, pt(0.5)));               // w0
            T[L::Z(i)][0] = iv2mt(isub(pt(1.0), u0));                      // z0
            const mt u1 = iv2mt(v0);                                                    // u[1] = v[0]/1
            T[L::U(i)][1] = u1;
            mt z1; z1.m = -u1.m; z1.t = u1.t;
            T[L::Z(i)][1] = z1;
        }
        for (int t = tid; t < N * D; t += NT) {
            const int i = t / D, cc = t % D;
            T[L::PU(i, cc)][0] = iv2mt(pt(cc == i ? 1.0 : 0.0));
            T[L::PV(i, cc)][0] = iv2mt(pt(cc == N + i ? 1.0 : 0.0));
            T[L::PU(i, cc)][1] = iv2mt(pt(cc == N + i ? 1.0 : 0.0));       // Pu[1] = Pv[0]/1
        }
        __syncthreads();
        PROF(0);
        TRC(2);
        table_loop<N>(S);                                         // Taylor tables of the centre trajectory (two orders per stage)
        bar_pair();                                               // both warps are done with the tables
        PROF(10);
        PROF(1);
        TRC(3);
        if (align_sm != ~0u && align_phase != 0) { if (tid == 0) sm_align_wait(align_sm); __syncthreads(); }      // the other half of an SM's fronts meets the same barrier half a step later
        // ---- step image: y = Phi(x), Jc = DPhi(x) (centre tables); frame at step start W = ((Tc + sum_j Theta_j r0_j) + R) + Ps Err ----
        //      One instruction stream per warp: the first threads evaluate TWO increment series each (interleaved Horner chains), the
        //      last D*N threads form the frame entries.
        {
            constexpr int NH = D + D * D, NHT = (NH + 1) / 2;
            for (int tt = tid; tt < NHT; tt += NT) {
                const int hA = image_series<N>(tt, 0), hB = image_series<N>(tt, 1);       // the two series of this lane (hB = NH: none)
                auto row_of = [&](int h) -> const mt* {
                    if (h < D) return h < N ? T[L::U(h)] : T[L::V(h - N)];
                    const int e = h - D, row = e / D, cc = e % D, i = row % N;
                    return row < N ? T[L::PU(i, cc)] : T[L::PV(i, cc)];
                };
                const mt* ra = row_of(hA); const mt* rb = row_of(hB < NH ? hB : NH - 1);
                ival a = mt2iv(ra[p]), bq = mt2iv(rb[p]);
#pragma unroll UNROLL_H
                for (int k = p - 1; k >= 1; k--) { a = horner_pt(a, hs, mt2iv(ra[k])); bq = horner_pt(bq, hs, mt2iv(rb[k])); }
                a = mk(__dmul_rd(a.lo, hs), __dmul_ru(a.hi, hs)); bq = mk(__dmul_rd(bq.lo, hs), __dmul_ru(bq.hi, hs));     // = eval_inc
                auto put = [&](int h, const ival& v) {
                    if (h < D) { S.y[h] = iadd(v, sym(mj.rho_phi)); }                                   // increment Phi(x) - x
                    else {
                        const int e = h - D, row = e / D, cc = e % D;
                        const ival kc = iadd(v, sym(mj.rho_psi));                                       // DPhi(x) - I
                        S.Kc[row][cc] = kc; S.Jc[row][cc] = iadd(pt(row == cc ? 1.0 : 0.0), kc);
                    }
                };
                put(hA, a);
                if (hB < NH) put(hB, bq);
            }
            for (int e = tid - (NT - (D * N) % NT) % NT; e < D * N; e += NT) {
                if (e < 0) continue;
                const int cc = e / N, bb = e % N;
                ival s0;
                const ival w = frame_entry<N>(S, cc, bb, &s0);
                S.S0[cc][bb] = s0; S.Ws[cc][bb] = w; S.Wm[cc][bb] = iv2mt(w);
                const ival rr = S.R[cc][bb];
                S.Rm[cc][bb] = iv2mt(rr); S.Tdm[cc][bb] = iv2mt(iadd(s0, rr)); S.Psm[cc][bb] = iv2mt(S.Ps[cc][bb]);
            }
        }
        __syncthreads();
        PROF(2);
        TRC(4);
        // ---- first / last frame data (topFrame::getFirstFrame / getLastFrame): rare ----
        if (s == 0) {
            for (int t = tid; t < D * N; t += NT) { const int rr = t / N, k = t % N; gres->first_frame[rr * (CCP_MAX_N + 1) + k] = toc(S.Ws[rr][k]); }
            if (tid < N && g > 0) {
                const double t1 = (1 < g) ? __ddiv_rd(__dmul_rd(1.0, hs), (double)g) : hs;
                ival du, dv; phi_prime<N>(S, Tf, tid, p, cur, hs, 0.0, t1, mj, du, dv);
                gres->first_frame[tid * (CCP_MAX_N + 1) + N] = toc(du);
                gres->first_frame[(N + tid) * (CCP_MAX_N + 1) + N] = toc(dv);
            }
        }
        if (last && tid < N && g > 0) {
            constexpr int ld = CCP_MAX_N + 2;
            const double tl = __ddiv_rd(__dmul_rd((double)(g - 1), hs), (double)g);
            ival du, dv; phi_prime<N>(S, Tf, tid, p, cur, hs, tl, hs, mj, du, dv);
            gres->last_frame[tid * ld + N] = toc(du);
            gres->last_frame[(N + tid) * ld + N] = toc(dv);
        }
        if ((s == 0 || last) && g > 0) __syncthreads();      // phi_prime read Psi^v rows that the frame tables are about to overlay
        // ---- frame tables F_k = Psi^u_k * Ws (the box for truncation and for Psi(phi) - Psi(x) is added to F_0 below) ----
        //      grid == 0: flow map only (boundaryValueProblem::Gxy / DGxy; include/ccp.h flow_P): no frame tables, no grid loop
        if (g > 0) for (int t = tid; t < (p + 1) * N * N; t += NT) {
            const int k = t / (N * N), a = (t / N) % N, bb = t % N;
            acc4 ac = acc_zero();
#pragma unroll
            for (int cc = 0; cc < D; cc++) acc_term(ac, T[L::PU(a, cc)][k], S.Wm[cc][bb]);
            S.fk()[k][a][bb] = acc_finish(ac);
        }
        __syncthreads();                            // Fk complete; all Taylor tables and un.scratch are dead from here (PT overlays them)
        PROF(3);
        TRC(5);
        second_variation<N>(S, PT, cur, hs, mj.g5, g > 0);          // Xi_j, dJ, J = Jc + [0,1] dJ, boxes added to F_0 / boxFd
        PROF(11);
        TRC(9);
        if (g > 0) {                                // grid == 0: flow map only, no grid loop / determinants / decisions
        // ---- grid loop (utils.cpp:219-252).  A lane owns ONE frame entry c and every (32/N^2)-th grid point: each coefficient
        //      Fk[k][c] is loaded once and feeds PPL independent Horner chains (shared-memory loads, not DFMAs, bound this loop).
        //      warp 0: point enclosures F(tau_i), i = 1..g   (F(0) = Fk[0] exactly)       -> EV[i]
        //      warp 1: range enclosures F([tau_i, tau_{i+1}]), i = 0..g-1                  -> EV[g+1+i]
        //      The derivative enclosures are formed on demand by the deciding lane (deriv_entries), as in the reference
        //      calculateDerivative is only reached when the rough test fails (topFrame.cpp:193-262).
        {
            constexpr int NN = N * N, NGRP = 32 / NN, PPL = (GMAX + NGRP - 1) / NGRP;
            const int c = lane % NN, grp = lane / NN;
            if (grp < NGRP) {
                const ival* q = &S.fk()[0][0][0] + p * NN + c;
                ival acc[PPL];
                double tl[PPL], th[PPL];
                const ival top = *q;
#pragma unroll
                for (int j = 0; j < PPL; j++) {
                    const int i = grp + j * NGRP, ic = i < g ? i : g - 1;       // point ic+1 (warp 0) / range [ic, ic+1] (warp 1)
                    acc[j] = top; tl[j] = S.gp[ic]; th[j] = S.gp[ic + 1];
                }
                if (wq == 0) {
#pragma unroll UNROLL_G
                    for (int k = p - 1; k >= 0; k--) {
                        q -= NN;
                        const ival ck = *q;
#pragma unroll
                        for (int j = 0; j < PPL; j++) acc[j] = horner_pt(acc[j], th[j], ck);
                    }
                } else {
#pragma unroll UNROLL_G
                    for (int k = p - 1; k >= 0; k--) {
                        q -= NN;
                        const ival ck = *q;
#pragma unroll
                        for (int j = 0; j < PPL; j++) acc[j] = horner_rg(acc[j], tl[j], th[j], ck);
                    }
                }
                const int base = (wq == 0 ? 1 : g + 1) * NN + c;
#pragma unroll
                for (int j = 0; j < PPL; j++) { const int i = grp + j * NGRP; if (i < g) EV[base + i * NN] = acc[j]; }
            }
        }
        __syncthreads();
        PROF(4);
        TRC(6);
        // ---- determinants (topFrame::constructDetSeries): one thread per point / range frame ----
        for (int t = tid; t < 2 * g + 1; t += NT) {
            ival Fm[N][N];
#pragma unroll
            for (int a = 0; a < N; a++)
#pragma unroll
                for (int bb = 0; bb < N; bb++) Fm[a][bb] = (t == 0) ? S.fk()[0][a][bb] : EV[(t * N + a) * N + bb];
            const ival dt = Det<N>::det(Fm);
            DT[t] = S.renorms ? imul_nl(dt, S.sc) : dt;          // determinant of the TRUE frame (renormalise)
        }
        __syncthreads();
        // ---- topFrame::countZeros decision per sub-interval (threads 0..g-1), tally by thread 0.  As in the reference the
        //      Jacobi derivative (calculateDerivative) is only formed when the rough enclosures do not settle the sub-interval ----
        if (wq == 0) {                              // warp 0: lane gi < g decides sub-interval gi; ballots replace a serial tally
            const int gi = lane;
            int decision = 0;
            ival timeI = pt(0.0);
            if (gi < g) {
                const double tl = S.gp[gi], th = S.gp[gi + 1];
                const ival tprev = S.st.tprev;
                timeI = mk(__dadd_rd(tprev.lo, tl), __dadd_ru(tprev.hi, th));
                const ival dL = DT[gi], dR = DT[gi + 1], dV = DT[g + 1 + gi];
                const ival LR = imul_nl(dL, dR);
                const bool need_D = (series != nullptr) || !(LR.lo > 0 && idefinite(dV));
                ival dD = pt(0.0);
                if (need_D) {
                    ival Fm[N][N], Dm[N][N];
#pragma unroll
                    for (int a = 0; a < N; a++)
#pragma unroll
                        for (int bb = 0; bb < N; bb++) Fm[a][bb] = EV[((g + 1 + gi) * N + a) * N + bb];
                    deriv_entries<N>(&S.fk()[0][0][0], &S.boxFd[0][0], p, tl, th, Dm);
                    dD = jacobi_derivative<N>(Fm, Dm);
                    if (S.renorms) dD = imul_nl(dD, S.sc);
                }
                decision = count_decide(timeI, dL, dR, dV, dD);
                if (!(ifinite(dL) && ifinite(dR) && ifinite(dV) && ifinite(dD))) decision |= 4;
                const int idx = s * g + gi;
                if (series != nullptr && idx < series_cap) {
                    ccp_series_rec rec; rec.time = toc(timeI); rec.det_L = toc(dL); rec.det_R = toc(dR); rec.det_V = toc(dV); rec.det_D = toc(dD);
                    series[idx] = rec;
                }
            }
            const unsigned FULLM = 0xffffffffu;
            const unsigned zmask = __ballot_sync(FULLM, (decision & 3) == 1);
            const unsigned fmask = __ballot_sync(FULLM, (decision & 3) == 2);
            const unsigned nmask = __ballot_sync(FULLM, (decision & 4) != 0);
            if (nmask && lane == 0) S.flag |= 1;
            const int n_conj = S.st.n_conj;                   // every lane reads the tally before lane 0 updates it
            __syncwarp();
            if (zmask && (decision & 3) == 1) {
                const int slot = n_conj + __popc(zmask & ((1u << lane) - 1u));
                if (slot < CCP_MAX_CONJ) { gres->conj_time[slot] = toc(timeI); gres->conj_index[slot] = s * g + gi; }
            }
            if (lane == 0) {
                if (fmask) { if (S.st.failure_count == 0) S.st.first_failure = s * g + (__ffs(fmask) - 1); S.st.failure_count += __popc(fmask); }
                if (zmask) { S.st.n_conj = n_conj + __popc(zmask); S.st.zero_count += __popc(zmask); }
                if (s == 0) S.st.det_first = DT[g + 1];
                if (last) S.st.det_last = DT[2 * g];
            }
        }
        }
        PROF(5);
        TRC(7);
        if (tid == 0) { const ival tp = S.st.tprev; S.st.tprev = mk(__dadd_rd(tp.lo, hs), __dadd_ru(tp.hi, hs)); }      // read again after the barriers of the set update
        if (set_update<N>(S, PT, cur, nxt)) status = CCP_ST_NONFINITE;
        PROF(6);
        TRC(8);
        if (g > 0 && !last && status == CCP_OK && ((s + 1) & S.cfg.renorm_mask) == 0) renormalise<N>(S);
        if (last && status == CCP_OK) {            // topFrame::getLastFrame: frame and front point at the end of the last step
            constexpr int ld = CCP_MAX_N + 2;
            for (int t = tid; t < D * N; t += NT) {
                const int rr = t / N, k = t % N;
                ival s0;
                ival wl = frame_entry<N>(S, rr, k, &s0);
                // C^1 flow map times A_u (boundaryValueProblem::DGxy): unstable columns (Tc + sum_j Theta_j r0_j) + R, stable columns Ps
                ival tu = iadd(iadd(pt(S.Tc[rr][k]), s0), S.R[rr][k]);
                if (S.renorms) {
                    // back to the basis A_u of the descriptor, piece by piece so that the affine dependence on r0 survives:
                    //   T = (Tc G) + sum_j (Theta_j G) r0_j + R G ;   W = T + Ps Err_0   (Err_0 = the descriptor's Eu_err; Err = Err_0 G^-1)
                    ival tc = pt(0.0), rg = pt(0.0), sj = pt(0.0);
                    for (int l = 0; l <= k; l++) ipfma(tc, S.G[l][k], S.Tc[rr][l]);
                    for (int j = 0; j < D; j++) {
                        ival th = pt(0.0);
                        for (int l = 0; l <= k; l++) ipfma(th, S.G[l][k], S.Th[j][rr][l]);
                        sj = iadd(sj, imul_nl(th, S.r0[j]));
                    }
                    for (int l = 0; l <= k; l++) rg = iadd(rg, imul_nl(S.R[rr][l], S.G[l][k]));
                    tu = iadd(iadd(tc, sj), rg);
                    acc4 e = acc_zero();
                    for (int j = 0; j < N; j++) acc_term(e, iv2mt(S.Ps[rr][j]), iv2mt(fromc(gdesc->Eu_err[j * CCP_MAX_N + k])));
                    wl = iadd(tu, acc_finish(e));
                }
                gres->last_frame[rr * ld + k] = toc(wl);
                gres->flow_P[rr * CCP_MAX_DIM + k] = toc(tu);
                gres->flow_P[rr * CCP_MAX_DIM + N + k] = toc(S.Ps[rr][k]);
            }
            if (tid < D) {
                ival a = pt(S.x[nxt][tid]);                       // nxt: the set after the update
                for (int j = 0; j < D; j++) ipfma(a, S.r0[j], S.C[nxt][tid][j]);          // fused: one rounding per bound and term
                gres->last_frame[tid * ld + N + 1] = toc(iadd(a, S.r[nxt][tid]));
                // flow mode: image of the CENTRE of the set (rho = 0 in x + C rho + r, valid for every rho in r0 pointwise): with flow_P it gives
                // the image of any sub-box of Uxy by the mean-value form (host/ccp_bvp.hpp: G at the point from the integration over the box)
                if (g == 0) gres->last_frame[tid * ld + N] = toc(iadd(pt(S.x[nxt][tid]), S.r[nxt][tid]));
            }
        }
    }
    return status;
}

// ------------------------------------------------------------------------------------------------
// One CTA (NT threads) = one front.  `series` (optional) receives steps*grid records.
// ------------------------------------------------------------------------------------------------
template <int N>
__device__ void run_front(FrontSmem<N>& S, const ccp_front_desc* __restrict__ gdesc, ccp_front_result* __restrict__ gres,
                          ccp_series_rec* __restrict__ series, int series_cap, unsigned align_sm, int align_phase) {
    using L = Lay<N>;
    constexpr int D = 2 * N;
    const int tid = threadIdx.x ^ S.vswap;
    mt (*T)[KMAX] = S.T;
    (void)sizeof(PhaseCheck<N>);

    // ---- stage the descriptor: coalesced 16-byte loads HBM -> shared ----
    {
        const int4* src = reinterpret_cast<const int4*>(gdesc);
        int4* dst = reinterpret_cast<int4*>(&T[0][0]);       // staged in the (still unused) table array
        for (int i = tid; i < (int)(sizeof(ccp_front_desc) / 16); i += NT) dst[i] = src[i];
    }
    __syncthreads();
    const ccp_front_desc& sdesc = *reinterpret_cast<const ccp_front_desc*>(&T[0][0]);
    const int p = sdesc.order, g = sdesc.grid, step_log2 = sdesc.step_log2;
    const double L_minus = sdesc.L_minus;
    int status = CCP_OK;
    if (sdesc.n != N || p < 2 || p > CCP_MAX_ORDER || g < 0 || g > GMAX || step_log2 < 1 || step_log2 > 20 || !(L_minus > 0) || !isfinite(L_minus) ||
        ldexp(L_minus, step_log2) > 16777216.0)          // more than 2^24 steps: the 32-bit step and sub-interval counters would overflow
        status = CCP_ST_BAD_DESC;
    if (status != CCP_OK) {
        __syncthreads();
        if (tid == 0) {
            ccp_front_result z; memset(&z, 0, sizeof(z)); z.status = status; z.first_failure = -1;
            *gres = z;
        }
        return;
    }
    const double h = ldexp(1.0, -step_log2);
    {
        const ccp_front_desc& d = sdesc;
        if (tid < N) S.bpar[tid] = fromc(d.b[tid]);
        if (tid < N - 1) S.cpar[tid] = fromc(d.c[tid]);
        if (tid < D) {
            ival pi = fromc(d.p_i[tid]);
            double xm = mid_ru(pi);
            S.x[0][tid] = xm; S.r[0][tid] = isub(pi, pt(xm)); S.r0[tid] = fromc(d.Uxy[tid]);
        }
        for (int t = tid; t < D * D; t += NT) { int i = t / D, j = t % D; double a = d.A_u[i * CCP_MAX_DIM + j]; S.C[0][i][j] = a; if (j < N) { S.Tc[i][j] = a; S.R[i][j] = pt(0.0); } else S.Ps[i][j - N] = pt(a); }
        for (int t = tid; t < D * D * N; t += NT) (&S.Th[0][0][0])[t] = 0.0;
        for (int t = tid; t < N * N; t += NT) { mt z; z.m = 0.0; z.t = 0.0; S.hm.M0[t / N][t % N] = z; S.hm.M1[t / N][t % N] = z; }
        if (tid >= 32 && tid < 32 + D + 1) { const int j = tid - 32; S.r0m[j] = j < D ? iv2mt(fromc(d.Uxy[j])) : ptm(1.0); }
        for (int t = tid; t < N * N; t += NT) { const int j = t / N, k = t % N; S.Errm[j][k] = iv2mt(fromc(d.Eu_err[j * CCP_MAX_N + k])); }
        for (int t = tid; t < N * N; t += NT) { int j = t / N, k = t % N; S.Err[j][k] = fromc(d.Eu_err[j * CCP_MAX_N + k]); }
    }
    for (int t = tid; t < (int)(sizeof(ccp_front_result) / 8); t += NT) reinterpret_cast<double*>(gres)[t] = 0.0;
    __syncthreads();                                // the staged descriptor is dead from here on
    if (tid == 0) { S.mj = majorant<N>(S.bpar, S.cpar, p, h); S.flag = 0; S.renorms = 0; S.sc = pt(1.0); }
    if (tid >= 32 && tid < 32 + N * N) { const int t = tid - 32; S.G[t / N][t % N] = pt(t / N == t % N ? 1.0 : 0.0); }
    if (tid == 32) majorant2<N>(S.bpar, S.cpar, h, S.g15);
    if (tid < N) {                                  // oracle/c2.hpp:make_par
        const int i = tid;
        mt z; z.m = 0.0; z.t = 0.0;
        const ival b = S.bpar[i];
        S.pm.B[i] = iv2mt(b); S.pm.B3[i] = iv2mt(imulk(b, 3)); S.pm.B6[i] = iv2mt(imulk(b, 6)); S.pm.halfB[i] = ihalf(b);
        S.pm.Cm[i] = i > 0 ? iv2mt(S.cpar[i > 0 ? i - 1 : 0]) : z;                 S.pm.Cp[i] = i < N - 1 ? iv2mt(S.cpar[i < N - 1 ? i : 0]) : z;
        S.pm.C2m[i] = i > 0 ? iv2mt(imulk(S.cpar[i > 0 ? i - 1 : 0], 2)) : z;      S.pm.C2p[i] = i < N - 1 ? iv2mt(imulk(S.cpar[i < N - 1 ? i : 0], 2)) : z;
        S.pm.C2[i] = S.pm.C2p[i];
        for (int l = 0; l < N; l++) for (int sl = 0; sl < 3; sl++) {
            mt v = z;
            if (i == l) v = sl == 0 ? S.pm.B6[i] : (sl == 1 ? S.pm.C2m[i] : S.pm.C2p[i]);
            else if (l == i + 1) v = sl < 2 ? S.pm.C2p[i] : z;                    // 2 c_i
            else if (l == i - 1) v = sl < 2 ? S.pm.C2m[i] : z;                    // 2 c_{i-1}
            S.pm.PN[i][l][sl] = v;
        }
    }
    __syncthreads();
    if (tid == 0) { S.mj.g1 = S.g15[0]; S.mj.g5 = S.g15[1]; S.st.zero_count = 0; S.st.failure_count = 0; S.st.first_failure = -1; S.st.n_conj = 0; S.st.tprev = pt(0.0); S.st.det_first = pt(0.0); S.st.det_last = pt(0.0); }
    __syncthreads();
    const Majorant& mj = S.mj;                              // stays in shared memory: a register copy would be live across the whole step loop
    if (!mj.ok) status = CCP_ST_ENCLOSURE;

    if (tid == 0) {
        S.cfg.p = p; S.cfg.g = g; S.cfg.step_log2 = step_log2; S.cfg.Stot = step_count(L_minus, h); S.cfg.renorm_mask = (1 << step_log2) - 1;      // basis renormalisation once per unit of time
        S.cfg.series_cap = series_cap; S.cfg.L_minus = L_minus; S.cfg.h = h; S.cfg.gdesc = gdesc; S.cfg.gres = gres; S.cfg.series = series;
    }
    __syncthreads();
    const int Stot = S.cfg.Stot;
    for (int s = 0; s < Stot && status == CCP_OK; s++) status = step_body<N>(S, s, align_sm, align_phase);
    __syncthreads();
    if (tid == 0) {
        const int n_conj = S.st.n_conj;
        gres->status = status; gres->zero_count = S.st.zero_count; gres->failure_count = S.st.failure_count;
        gres->series_length = Stot * g; gres->steps = Stot; gres->n_conj = n_conj < CCP_MAX_CONJ ? n_conj : CCP_MAX_CONJ;
        gres->first_failure = S.st.first_failure; gres->reserved = 0;
        gres->det_first = toc(S.st.det_first); gres->det_last = toc(S.st.det_last);
    }
}

#ifndef CCP_MINB
#define CCP_MINB 8
#endif
#ifdef CCP_MAXREG
#define CCP_KERNEL_ATTR __maxnreg__(CCP_MAXREG)
#else
#define CCP_KERNEL_ATTR __launch_bounds__(NT, (N == 4 ? 4 : CCP_MINB))
#endif
template <int N>
__global__ void CCP_KERNEL_ATTR front_kernel(const ccp_front_desc* __restrict__ descs, ccp_front_result* __restrict__ results,
                                                      const int* __restrict__ index, int count, ccp_series_rec* series, int series_cap, int flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrontSmem<N>& S = *reinterpret_cast<FrontSmem<N>*>(smem_raw);
    const unsigned align_sm = ((flags & 1) && series == nullptr && count > 1) ? sm_id() : ~0u;      // single fronts (plots) run alone
    __shared__ int s_ord;
    if (threadIdx.x == 0) s_ord = align_sm != ~0u ? (int)sm_align_join(align_sm) : 0;
    __syncthreads();
    // The warps of a CTA sit on neighbouring SM sub-partitions (hardware warp slot mod 4), so with fixed roles every Psi warp of an SM (the heavier
    // one in the table loop) would share two of the four FP64 pipes.  Every second CTA of a sub-partition pair exchanges the roles of its warps.
    if (threadIdx.x == 0) { unsigned wid; asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid)); S.vswap = ((flags & 16) && ((wid >> 2) & 1u)) ? 32 : 0; }
    __syncthreads();
    const int pm = (flags >> 1) & 7;                          // experiment: which bit of the per-SM ordinal selects the half-step offset (0 = none)
    const int align_phase = pm ? (s_ord >> (pm - 1)) & 1 : 0;
    for (int w = blockIdx.x; w < count; w += gridDim.x) {
        const int f = index ? index[w] : w;
        run_front<N>(S, descs + f, results + f, series, series_cap, align_sm, align_phase);
        __syncthreads();
    }
    if (align_sm != ~0u && threadIdx.x == 0) sm_align_leave(align_sm);
}

} // namespace ccp
