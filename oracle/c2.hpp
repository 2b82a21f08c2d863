// oracle/c2.hpp -- TEST INFRASTRUCTURE (CPU oracle, mode "c2").  Not part of the product path.
//
// Mode c1 plus FIRST-ORDER TRACKING of the dependence of the frame on the initial box of the front
// (the reference gets this from the off-diagonal block of the 4n x 4n Jacobian of f_linearize that
// CAPD's C0Rect2Set carries along, propagateManifold.cpp:110-114).
//
//   phi-set     x + C r0 + r                                   (as c1)
//   Psi tables  at the CENTRE x (thin), order p                (c1: over the hull)
//   step Jacobian over the set, mean-value form in the r0-directions c_j = C e_j:
//       J(phi) in Jc + sum_j Xi_j[X] r0_j + E_r ,   Xi_j[X] = enclosure over the hull of D(DPhi_h)[c_j]
//     Xi_j is the second variation; with N(t) = DM(u(t))[Psi^u(t) c_j] its Taylor expansion to h^3 is closed:
//       Xi^u = [ h^2 N0/2 + h^3 N1/6                     | h^3 N0/6          ]
//       Xi^v = [ h N0 + h^2 N1/2 + h^3 (N2 + (M0 N0 + N0 M0)/2)/3 | h^2 N0/2 + h^3 N1/3 ]
//     plus a majorant remainder gamma_4 h^4 |c_j|_1 per entry (c2_majorant).
//   unstable columns of P = Dphi_t A_u:   T = Tc + sum_j Theta_j r0_j + R   (Tc, Theta_j points, R interval)
//       T+ = J(phi) T  =>  Tc+ = mid(Jc Tc),  Theta_j+ = mid(Jc Theta_j + Xi_j Tc),  R+ = everything else
//   stable columns Ps (they only enter through Err ~ 1e-5): interval matrix, Ps+ = J[X] Ps  (as c1)
//   frame  W = T + Ps Err ;  frame tables F_k = Psi_c^u_k W + one box for (Psi(phi) - Psi(x)) W.
#pragma once
#include "c1.hpp"

namespace orc {
namespace c2 {

using namespace c1;

struct Maj2 { double g1, g5; };     // per unit |c|_1: |Xi(tau)| <= g1 tau ;  |Xi(tau) - sum_{k<=4} Xi_k tau^k| <= g5 tau^5

// Scalar majorant of the second variation xi' = Df xi + D^2f[Psi c, psi] on the a-priori box G (see c1_majorant
// for w, g, Mh, be).  |N(d)|_inf <= 6 (bh+cs) |w| |d|,  |d| <= beta |c|_1.
static inline Maj2 c2_majorant(int n, const ival* b, const ival* c, double h) {
    double bh = 0, cs = 0;
    for (int i = 0; i < n; i++) bh = std::max(bh, mag(b[i]));
    for (int i = 0; i < n; i++) { double s = 0; if (i > 0) s = add_ru(s, mag(c[i - 1])); if (i < n - 1) s = add_ru(s, mag(c[i])); cs = std::max(cs, s); }
    const double W0 = 0.5625, G0 = 0.25, V0 = 0.5;
    const int K = 5;
    double bc = add_ru(bh, cs);
    double F0 = mul_ru(mul_ru(bc, G0), W0);
    double w[K + 2], g[K + 2], be[K + 2], Mh[K + 2], ga[K + 2];
    w[0] = W0; g[0] = G0; w[1] = std::max(V0, F0);
    for (int k = 1; k <= K; k++) {
        double s = 0; for (int j = 0; j <= k; j++) s = fma_ru(w[j], w[k - j], s);
        g[k] = s;
        double f = 0; for (int j = 0; j <= k; j++) f = fma_ru(g[j], w[k - j], f);
        f = mul_ru(bc, f);
        w[k + 1] = div_ru(std::max(w[k], f), (double)(k + 1));
    }
    double M0 = add_ru(add_ru(mul_ru(bh, 0.75), mul_ru(cs, 0.25)), mul_ru(mul_ru(2.0, cs), mul_ru(W0, W0)));
    double a = std::max(1.0, M0), ah = mul_ru(a, h);
    be[0] = add_ru(add_ru(1.0, ah), mul_ru(ah, ah));
    Mh[0] = M0;
    double k3 = mul_ru(3.0, bc);
    for (int k = 1; k <= K; k++) Mh[k] = mul_ru(k3, g[k]);
    for (int k = 0; k <= K; k++) {
        double s = 0; for (int j = 0; j <= k; j++) s = fma_ru(Mh[j], be[k - j], s);
        be[k + 1] = div_ru(std::max(be[k], s), (double)(k + 1));
    }
    const double H6 = mul_ru(6.0, bc);
    // a-priori bound of |xi(tau)|, tau in [0,h]:  tau e^{a tau} (6 bc W0) beta0^2 <= h beta0^3 6 bc W0
    ga[0] = mul_ru(mul_ru(h, mul_ru(be[0], mul_ru(be[0], be[0]))), mul_ru(H6, W0));
    double bb[K + 2];                                       // (beta*beta)_k
    for (int k = 0; k <= K; k++) { double s = 0; for (int j = 0; j <= k; j++) s = fma_ru(be[j], be[k - j], s); bb[k] = s; }
    for (int k = 0; k < K; k++) {
        double s = 0; for (int j = 0; j <= k; j++) s = fma_ru(Mh[j], ga[k - j], s);
        double f = 0; for (int j = 0; j <= k; j++) f = fma_ru(w[j], bb[k - j], f);
        s = fma_ru(H6, f, s);
        ga[k + 1] = div_ru(std::max(ga[k], s), (double)(k + 1));
    }
    Maj2 m; m.g1 = ga[1]; m.g5 = ga[5];
    return m;
}


static inline ival half(const ival& a) { return ival(0.5 * a.lo, 0.5 * a.hi); }
// acc*h + c for a point h > 0
static inline ival horner_hp(const ival& acc, double h, const ival& c) { return ival(fma_rd(acc.lo, h, c.lo), fma_ru(acc.hi, h, c.hi)); }
// a += x * p for a point p: one directed FMA per bound (tighter than a rounded product plus a rounded sum)
static inline void ipfma(ival& a, const ival& x, double p) {
    const double l = p >= 0 ? x.lo : x.hi, h = p >= 0 ? x.hi : x.lo;
    a.lo = fma_rd(l, p, a.lo); a.hi = fma_ru(h, p, a.hi);
}
// sum_{k>=1} s_k tau^k by Horner (the order-0 coefficient is left out: increments keep their own, much finer, rounding grid)
static inline ival eval_inc(const mt* s, int p, double tau) {
    ival acc = mt2iv(s[p]);
    for (int k = p - 1; k >= 1; k--) acc = horner_pt(acc, tau, mt2iv(s[k]));
    return ival(mul_rd(acc.lo, tau), mul_ru(acc.hi, tau));
}
static inline mt ptm(double v) { mt o; o.m = v; o.t = std::fabs(v); return o; }      // iv2mt of a point
static const mt MT0 = {0.0, 0.0};

// Parameter tables of a front in (mid,t) form; missing chain neighbours are zero parameters so that every row has the
// same three terms (i, i-1, i+1 with clamped indices): the GPU evaluates these sums branch-free.
struct ParM {
    mt B[MAXN], B3[MAXN], B6[MAXN], Cm[MAXN], Cp[MAXN], C2m[MAXN], C2p[MAXN], C2[MAXN];
    ival halfB[MAXN];
    mt PN[MAXN][MAXN][3];          // N_k[i][l] = sum_s PN[i][l][s] * WD[a_s][b_s]   (slot indices: slot_ab)
};
static inline void make_par(ParM& P, int n, const ival* b, const ival* c) {
    for (int i = 0; i < n; i++) {
        P.B[i] = iv2mt(b[i]); P.B3[i] = iv2mt(mulk(b[i], 3)); P.B6[i] = iv2mt(mulk(b[i], 6)); P.halfB[i] = half(b[i]);
        P.Cm[i] = i > 0 ? iv2mt(c[i - 1]) : MT0;                 P.Cp[i] = i < n - 1 ? iv2mt(c[i]) : MT0;
        P.C2m[i] = i > 0 ? iv2mt(mulk(c[i - 1], 2)) : MT0;       P.C2p[i] = i < n - 1 ? iv2mt(mulk(c[i], 2)) : MT0;
        P.C2[i] = i < n - 1 ? iv2mt(mulk(c[i], 2)) : MT0;
    }
    for (int i = 0; i < n; i++) for (int l = 0; l < n; l++) for (int s = 0; s < 3; s++) {
        mt v = MT0;
        if (i == l) v = s == 0 ? P.B6[i] : (s == 1 ? P.C2m[i] : P.C2p[i]);
        else if (l == i + 1 || l == i - 1) v = s < 2 ? P.C2[i < l ? i : l] : MT0;
        P.PN[i][l][s] = v;
    }
}
// indices (a,b) of the product series w_a d_b that slot s of entry (i,l) of N refers to (clamped; unused slots carry a zero parameter)
static inline void slot_ab(int n, int i, int l, int s, int& a, int& b) {
    if (i == l) { a = b = s == 0 ? i : (s == 1 ? (i > 0 ? i - 1 : 0) : (i < n - 1 ? i + 1 : n - 1)); }
    else { const int e = i < l ? i : l; const int e1 = e + 1 < n ? e + 1 : n - 1; if (s == 0) { a = e1; b = e; } else { a = e; b = e1; } }
}

// low-order Taylor data over the hull [X]: W[q] = coefficients of u - 1/2 (q <= 3), G = u(1-u), dense tridiagonal M0, M1
struct HullM { mt W[4][MAXN], G[MAXN], Pw[MAXN], M0[MAXN][MAXN], M1[MAXN][MAXN]; };
static inline void hull_data(HullM& H, const ParM& P, int n, const ival* X) {
    ival w0[MAXN];
    for (int m = 0; m < n; m++) {
        w0[m] = X[m] - ival(0.5);
        H.W[0][m] = iv2mt(w0[m]); H.W[1][m] = iv2mt(X[n + m]); H.G[m] = iv2mt(X[m] * (ival(1.0) - X[m]));
        acc4 a; acc_term(a, H.W[0][m], H.W[1][m]); H.Pw[m] = iv2mt(acc_finish(a));
    }
    for (int i = 0; i < n; i++) for (int l = 0; l < n; l++) { H.M0[i][l] = MT0; H.M1[i][l] = MT0; }
    for (int e = 0; e + 1 < n; e++) {
        acc4 a; acc_term(a, H.W[0][e], H.W[0][e + 1]);
        acc4 a2; acc_term(a2, P.C2[e], iv2mt(acc_finish(a)));
        H.M0[e][e + 1] = H.M0[e + 1][e] = iv2mt(acc_finish(a2));
        acc4 u; acc_term(u, H.W[0][e], H.W[1][e + 1]); acc_term(u, H.W[1][e], H.W[0][e + 1]);
        acc4 u2; acc_term(u2, P.C2[e], iv2mt(acc_finish(u)));
        H.M1[e][e + 1] = H.M1[e + 1][e] = iv2mt(acc_finish(u2));
    }
    for (int i = 0; i < n; i++) {
        const int im = i > 0 ? i - 1 : 0, ip = i < n - 1 ? i + 1 : n - 1;
        acc4 sa; acc_term(sa, P.B[i], H.G[i]); acc_term(sa, P.Cm[i], H.G[im]); acc_term(sa, P.Cp[i], H.G[ip]);
        H.W[2][i] = iv2mt(half(-(w0[i] * acc_finish(sa))));                                   // u_2 = F(u)/2
        acc4 ta; acc_term(ta, P.B3[i], H.G[i]); acc_term(ta, P.Cm[i], H.G[im]); acc_term(ta, P.Cp[i], H.G[ip]);
        H.M0[i][i] = iv2mt(P.halfB[i] - acc_finish(ta));
        acc4 ma; acc_term(ma, P.B6[i], H.Pw[i]); acc_term(ma, P.C2m[i], H.Pw[im]); acc_term(ma, P.C2p[i], H.Pw[ip]);
        H.M1[i][i] = iv2mt(acc_finish(ma));
    }
    for (int i = 0; i < n; i++) {
        acc4 a; for (int l = 0; l < n; l++) acc_term(a, H.M0[i][l], H.W[1][l]);
        H.W[3][i] = iv2mt(divk(acc_finish(a), 6));                                            // u_3 = M0 u_1 / 6
    }
}

// S0 = sum_j Theta_j r0_j ;  W = ((Tc + S0) + R) + Ps Err
static inline void frame_of_state(int n, const double (*Tc)[MAXN], const double (*Th)[MAXD][MAXN], const ival (*R)[MAXN], const ival (*Ps)[MAXN],
                                  const ival* r0, const mt (*Errm)[MAXN], ival (*S0)[MAXN], ival (*W)[MAXN]) {
    const int D = 2 * n;
    for (int cc = 0; cc < D; cc++) for (int bb = 0; bb < n; bb++) {
        ival a(0.0); for (int j = 0; j < D; j++) ipfma(a, r0[j], Th[j][cc][bb]);
        S0[cc][bb] = a;
        acc4 e; for (int j = 0; j < n; j++) acc_term(e, iv2mt(Ps[cc][j]), Errm[j][bb]);
        W[cc][bb] = ((ival(Tc[cc][bb]) + S0[cc][bb]) + R[cc][bb]) + acc_finish(e);
    }
}

// ---- basis renormalisation of the frame (once per unit of time) ----
// The columns of T = Dphi_t A_u all align with the fastest unstable direction (condition number ~1e5 at L_-), so the interval
// determinant of the hull of T loses every digit as soon as the relative width of the entries (box of the front, amplified by
// e^{lambda t}) exceeds 1/cond -- with the validated neighbourhood of the connecting orbit (1e-13 .. 1e-12) that is the last third
// of the front.  The frame is only a BASIS of E^u: for any invertible point matrix H, det(top(T)) = det(top(T H)) / det(H), zeros and
// (for det H > 0) signs unchanged.  H = (R factor of a floating-point Gram-Schmidt of Tc)^-1 is computed in round-to-nearest
// (any H is valid; only determinism matters) and applied rigorously to every piece of the frame representation:
//     Tc <- mid([Tc H]),  Theta_j <- mid([Theta_j H]),  R <- R H + residuals,  Err <- Err H   (Ps Err H = Ps (Err H)).
// G encloses the product of the H^-1 (true frame = normalised frame * G, used for getLastFrame), sc encloses det(G).
static inline bool renorm_matrix(int n, const double (*Tc)[MAXN], double (*H)[MAXN]) {
    RoundNearestGuard rn;
    const int D = 2 * n;
    double Q[MAXD][MAXN], Rt[MAXN][MAXN];
    for (int k = 0; k < n; k++) {
        double v[MAXD];
        for (int i = 0; i < D; i++) v[i] = Tc[i][k];
        for (int m = 0; m < k; m++) {
            double dp = 0.0;
            for (int i = 0; i < D; i++) dp = dp + Q[i][m] * v[i];
            Rt[m][k] = dp;
            for (int i = 0; i < D; i++) v[i] = v[i] - dp * Q[i][m];
        }
        double nn = 0.0;
        for (int i = 0; i < D; i++) nn = nn + v[i] * v[i];
        nn = std::sqrt(nn);
        if (!(nn > 0.0) || !std::isfinite(nn)) return false;
        Rt[k][k] = nn;
        for (int i = 0; i < D; i++) Q[i][k] = v[i] / nn;
    }
    for (int k = 0; k < n; k++) {                         // H = Rt^-1, upper triangular, positive diagonal
        for (int i = 0; i < n; i++) H[i][k] = 0.0;
        H[k][k] = 1.0 / Rt[k][k];
        for (int i = k - 1; i >= 0; i--) {
            double a = 0.0;
            for (int m = i + 1; m <= k; m++) a = a + Rt[i][m] * H[m][k];
            H[i][k] = (-a) / Rt[i][i];
        }
    }
    for (int k = 0; k < n; k++) for (int i = 0; i <= k; i++) if (!std::isfinite(H[i][k])) return false;
    for (int k = 0; k < n; k++) if (!(H[k][k] > 0.0)) return false;
    return true;
}

static inline void run_front(const ccp_front_desc& d, ccp_front_result& res, std::vector<ccp_series_rec>* series) {
    std::memset(&res, 0, sizeof(res));
    res.first_failure = -1;
    const int n = d.n, D = 2 * n, p = d.order, g = d.grid;
    if (n < 2 || n > MAXN || p < 2 || p > MAXP || g < 0 || g > MAXG || d.step_log2 < 1 || d.step_log2 > 20 || !(d.L_minus > 0) || !std::isfinite(d.L_minus) || std::ldexp(d.L_minus, d.step_log2) > 16777216.0) {
        res.status = CCP_ST_BAD_DESC; return;
    }
    const double h = std::ldexp(1.0, -d.step_log2);
    ival b[MAXN], c[MAXN];
    for (int i = 0; i < n; i++) b[i] = fromc(d.b[i]);
    for (int i = 0; i < n - 1; i++) c[i] = fromc(d.c[i]);
    const int S = step_count(d.L_minus, h);
    res.steps = S; res.series_length = S * g;
    Majorant mj = c1_majorant(n, b, c, p, h);
    if (!mj.ok) { res.status = CCP_ST_ENCLOSURE; return; }
    const Maj2 m2 = c2_majorant(n, b, c, h);
    static thread_local ParM PM;
    make_par(PM, n, b, c);

    double x[MAXD], C[MAXD][MAXD];
    ival r0[MAXD], r[MAXD], Err[MAXN][MAXN];
    double Tc[MAXD][MAXN], Th[MAXD][MAXD][MAXN];           // Th[j] = Theta_j
    ival R[MAXD][MAXN], Ps[MAXD][MAXN];
    for (int i = 0; i < D; i++) {
        ival pi = fromc(d.p_i[i]);
        x[i] = mid_ru(pi); r[i] = pi - ival(x[i]); r0[i] = fromc(d.Uxy[i]);
        for (int j = 0; j < D; j++) C[i][j] = d.A_u[i * CCP_MAX_DIM + j];
        for (int k = 0; k < n; k++) { Tc[i][k] = C[i][k]; Ps[i][k] = ival(C[i][n + k]); R[i][k] = ival(0.0); }
        for (int j = 0; j < D; j++) for (int k = 0; k < n; k++) Th[j][i][k] = 0.0;
    }
    for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) Err[j][k] = fromc(d.Eu_err[j * CCP_MAX_N + k]);
    mt r0m[MAXD + 1], Errm[MAXN][MAXN];                      // (mid,t) forms; r0m[D] = 1 weights the direction "box r"
    for (int j = 0; j < D; j++) r0m[j] = iv2mt(r0[j]);
    r0m[D] = ptm(1.0);
    for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) Errm[j][k] = iv2mt(Err[j][k]);

    ival G[MAXN][MAXN], sc(1.0);                              // true frame = normalised frame * G ;  sc encloses det G > 0
    for (int i = 0; i < n; i++) for (int k = 0; k < n; k++) G[i][k] = ival(i == k ? 1.0 : 0.0);
    int renorms = 0;
    const int renorm_mask = (1 << d.step_log2) - 1;         // once per unit of time
    Counter cnt;
    ival tprev(0.0);
    static thread_local PhiTables ct;
    static thread_local PsiTables ps;
    ival Fk[MAXP + 1][MAXN][MAXN];
    ival FRprev[MAXN][MAXN];

    for (int s = 0; s < S; s++) {
        const bool last = (s == S - 1);
        const double hs = last ? sub_ru(d.L_minus, tprev.hi) : h;
        ival X[MAXD];
        for (int i = 0; i < D; i++) {
            ival a(x[i]);
            for (int j = 0; j < D; j++) ipfma(a, r0[j], C[i][j]);
            X[i] = a + r[i];
        }
        bool inside = true, fin = true;
        for (int i = 0; i < n; i++) {
            inside = inside && subset(X[i], mj.Gs_u) && subset(X[n + i], mj.Gs_v);
            fin = fin && finite(X[i]) && finite(X[n + i]);
        }
        if (!fin) { res.status = CCP_ST_NONFINITE; break; }
        if (!inside) { res.status = CCP_ST_ENCLOSURE; break; }
        // ---- centre tables (order p): phi and Psi along the centre trajectory ----
        ival xc[MAXD]; for (int i = 0; i < D; i++) xc[i] = ival(x[i]);
        phi_tables(ct, n, p, b, c, xc, xc + n);
        psi_tables(ps, ct, n, p, b, c);
        // ---- low-order Taylor data over the hull [X] ----
        static thread_local HullM H;
        hull_data(H, PM, n, X);
        double rnorm = 0; for (int i = 0; i < D; i++) rnorm = add_ru(rnorm, mag(r[i]));
        const double hs2 = mul_ru(hs, hs), hs5 = mul_ru(mul_ru(hs2, hs2), hs);
        // ---- second variation over the hull: directions c_j (weight r0_j, tracked), j < D, and the box r (weight 1) ----
        static thread_local ival Xi[MAXD + 1][MAXD][MAXD];     // Xi[j][row][col]
        static thread_local mt XiM[MAXD + 1][MAXD][MAXD];
        static thread_local double AM[MAXD + 1][MAXD][MAXD];
        for (int j = 0; j <= D; j++) {
            mt dd[4][MAXN];
            double c1n = 0;
            if (j < D) { for (int i = 0; i < D; i++) c1n = add_ru(c1n, std::fabs(C[i][j])); for (int i = 0; i < n; i++) { dd[0][i] = ptm(C[i][j]); dd[1][i] = ptm(C[n + i][j]); } }
            else       { c1n = rnorm; for (int i = 0; i < n; i++) { dd[0][i] = iv2mt(r[i]); dd[1][i] = iv2mt(r[n + i]); } }
            for (int m = 0; m < n; m++) {
                acc4 a2; for (int l = 0; l < n; l++) acc_term(a2, H.M0[m][l], dd[0][l]);
                dd[2][m] = iv2mt(half(acc_finish(a2)));
                acc4 a3; for (int l = 0; l < n; l++) acc_term(a3, H.M1[m][l], dd[0][l]);
                for (int l = 0; l < n; l++) acc_term(a3, H.M0[m][l], dd[1][l]);
                dd[3][m] = iv2mt(divk(acc_finish(a3), 6));
            }
            mt WD[4][MAXN][MAXN];                                  // (w_i d_m)_k
            for (int k = 0; k < 4; k++) for (int i = 0; i < n; i++) for (int m = 0; m < n; m++) {
                acc4 a; for (int q = 0; q <= k; q++) acc_term(a, H.W[q][i], dd[k - q][m]);
                WD[k][i][m] = iv2mt(acc_finish(a));
            }
            mt NN[4][MAXN][MAXN];                                  // Taylor coefficients of N = DM(u)[Psi^u c_j], dense
            for (int k = 0; k < 4; k++) for (int i = 0; i < n; i++) for (int l = 0; l < n; l++) {
                acc4 a;
                for (int sl = 0; sl < 3; sl++) { int ia, ib; slot_ab(n, i, l, sl, ia, ib); acc_term(a, PM.PN[i][l][sl], WD[k][ia][ib]); }
                NN[k][i][l] = iv2mt(acc_finish(a));
            }
            const double rem = mul_ru(mul_ru(m2.g5, hs5), c1n);
            for (int i = 0; i < n; i++) for (int k = 0; k < n; k++) {
                acc4 p00, p01, p10;
                for (int l = 0; l < n; l++) acc_term(p00, H.M0[i][l], NN[0][l][k]);
                for (int l = 0; l < n; l++) acc_term(p00, NN[0][i][l], H.M0[l][k]);
                for (int l = 0; l < n; l++) acc_term(p01, H.M0[i][l], NN[1][l][k]);
                for (int l = 0; l < n; l++) acc_term(p01, NN[0][i][l], H.M1[l][k]);
                for (int l = 0; l < n; l++) acc_term(p10, H.M1[i][l], NN[0][l][k]);
                for (int l = 0; l < n; l++) acc_term(p10, NN[1][i][l], H.M0[l][k]);
                const ival P00 = acc_finish(p00), P01 = acc_finish(p01), P10 = acc_finish(p10);
                const ival N0 = mt2iv(NN[0][i][k]), N1 = mt2iv(NN[1][i][k]), N2 = mt2iv(NN[2][i][k]), N3 = mt2iv(NN[3][i][k]);
                ival cf[4][4];                                    // cf[block][order-1]: blocks 0 UU, 1 UV, 2 VU, 3 VV
                const ival X3 = divk(N2 + half(P00), 3);          // Xi^v_3, u columns
                cf[0][0] = ival(0.0); cf[0][1] = half(N0); cf[0][2] = divk(N1, 6); cf[0][3] = divk(X3, 4);
                cf[1][0] = ival(0.0); cf[1][1] = ival(0.0); cf[1][2] = divk(N0, 6); cf[1][3] = divk(N1, 12);
                cf[2][0] = N0; cf[2][1] = half(N1); cf[2][2] = X3; cf[2][3] = divk(N3 + divk(P01, 6) + half(P10), 4);
                cf[3][0] = ival(0.0); cf[3][1] = half(N0); cf[3][2] = divk(N1, 3); cf[3][3] = divk(N2 + divk(P00, 6), 4);
                for (int q = 0; q < 4; q++) {
                    ival acc = cf[q][3]; double am = mag(cf[q][3]);
                    for (int e = 2; e >= 0; e--) { acc = horner_hp(acc, hs, cf[q][e]); am = fma_ru(am, hs, mag(cf[q][e])); }
                    acc = horner_hp(acc, hs, ival(0.0)); am = mul_ru(am, hs);
                    const int row = (q >> 1) * n + i, col = (q & 1) * n + k;
                    Xi[j][row][col] = acc + sym(rem);
                    XiM[j][row][col] = iv2mt(Xi[j][row][col]);
                    AM[j][row][col] = add_ru(am, rem);
                }
            }
        }
        ival dJ[MAXD][MAXD];                                    // sum_j Xi_j r0_j + Xi_D  : encloses DPhi(phi) - DPhi(x) on the set
        mt dJm[MAXD][MAXD];
        double magXu[MAXN][MAXD], magXv[MAXN][MAXD];            // sup over tau in [0,hs] of |Psi(tau;phi) - Psi(tau;x)|, u rows / v rows
        for (int row = 0; row < D; row++) for (int col = 0; col < D; col++) {
            acc4 a; double mg = 0.0;
            for (int j = 0; j <= D; j++) { acc_term(a, XiM[j][row][col], r0m[j]); mg = fma_ru(AM[j][row][col], j < D ? mag(r0[j]) : 1.0, mg); }
            dJ[row][col] = acc_finish(a); dJm[row][col] = iv2mt(dJ[row][col]);
            if (row < n) magXu[row][col] = mg; else magXv[row - n][col] = mg;
        }
        // ---- frame at step start:  W = (Tc + sum_j Theta_j r0_j + R) + Ps Err ----
        ival S0[MAXD][MAXN], Ws[MAXD][MAXN]; mt Wm[MAXD][MAXN];
        frame_of_state(n, Tc, Th, R, Ps, r0, Errm, S0, Ws);
        for (int cc = 0; cc < D; cc++) for (int bb = 0; bb < n; bb++) Wm[cc][bb] = iv2mt(Ws[cc][bb]);
        // truncation box of the centre tables and box for (Psi(tau;phi) - Psi(tau;x)) W, tau in [0,hs]; same for the derivative
        double boxF[MAXN][MAXN], boxFd[MAXN][MAXN];
        for (int bb = 0; bb < n; bb++) {
            double sm = 0; for (int cc = 0; cc < D; cc++) sm = add_ru(sm, mag(Ws[cc][bb]));
            const double rhoF = mul_ru(mj.rho_psi, sm), rhoFd = mul_ru(mj.rho_psi_d, sm);
            for (int a = 0; a < n; a++) {
                double s0 = rhoF, s1 = rhoFd;
                for (int cc = 0; cc < D; cc++) { s0 = fma_ru(magXu[a][cc], mag(Ws[cc][bb]), s0); s1 = fma_ru(magXv[a][cc], mag(Ws[cc][bb]), s1); }
                boxF[a][bb] = s0; boxFd[a][bb] = s1;
            }
        }
        // grid == 0: flow map only (boundaryValueProblem::Gxy / DGxy, include/ccp.h flow_P) -- no frame tables, no grid loop
        if (g > 0) for (int k = 0; k <= p; k++) for (int a = 0; a < n; a++) for (int bb = 0; bb < n; bb++) {
            acc4 ac; for (int cc = 0; cc < D; cc++) acc_term(ac, ps.Pu[a][cc][k], Wm[cc][bb]);
            ival f = acc_finish(ac);
            if (k == 0) f = f + sym(boxF[a][bb]);
            Fk[k][a][bb] = f;
        }
        // ---- grid loop + topFrame (as c1) ----
        double gp[MAXG + 1]; grid_points(hs, g, gp); if (g == 0) gp[1] = hs;
        Frame FL, FR, FV, FD;
        for (int i = 0; i < g; i++) {
            const double tl = gp[i], th = gp[i + 1];
            for (int a = 0; a < n; a++) for (int bb = 0; bb < n; bb++) {
                if (i == 0) {
                    ival acc = Fk[p][a][bb];
                    for (int k = p - 1; k >= 0; k--) acc = horner_pt(acc, tl, Fk[k][a][bb]);
                    FL[a][bb] = acc;
                } else FL[a][bb] = FRprev[a][bb];
                ival acc = Fk[p][a][bb];
                for (int k = p - 1; k >= 0; k--) acc = horner_pt(acc, th, Fk[k][a][bb]);
                FR[a][bb] = acc; FRprev[a][bb] = acc;
                acc = Fk[p][a][bb];
                for (int k = p - 1; k >= 0; k--) acc = horner_rg(acc, tl, th, Fk[k][a][bb]);
                FV[a][bb] = acc;
                acc = mulk(Fk[p][a][bb], p);
                for (int k = p - 1; k >= 1; k--) acc = horner_rg(acc, tl, th, mulk(Fk[k][a][bb], k));
                FD[a][bb] = acc + sym(boxFd[a][bb]);
            }
            ival time(add_rd(tprev.lo, tl), add_ru(tprev.hi, th));
            ival dL = detN(FL, n), dR = detN(FR, n), dV = detN(FV, n);
            if (renorms) { dL = dL * sc; dR = dR * sc; dV = dV * sc; }              // determinants of the TRUE frame
            const ival LR = dL * dR;
            const bool need_D = (series != nullptr) || !(LR.lo > 0 && definite(dV));
            ival dD = need_D ? jacobi_derivative(FV, FD, n) : ival(0.0);
            if (renorms && need_D) dD = dD * sc;
            int idx = s * g + i;
            count_step(cnt, idx, time, dL, dR, dV, dD);
            if (series) { ccp_series_rec rec; rec.time = toc(time); rec.det_L = toc(dL); rec.det_R = toc(dR); rec.det_V = toc(dV); rec.det_D = toc(dD); series->push_back(rec); }
            if (idx == 0) res.det_first = toc(dV);
            if (idx == S * g - 1) res.det_last = toc(dV);
            if (!(finite(dL) && finite(dR) && finite(dV) && finite(dD))) res.status = CCP_ST_NONFINITE;
        }
        // ---- step: y = Phi(x), Jc = DPhi(x) ----
        ival y[MAXD], Jc[MAXD][MAXD], Kc[MAXD][MAXD], J[MAXD][MAXD];
        for (int i = 0; i < n; i++) {
            y[i] = eval_inc(ct.u[i], p, hs) + sym(mj.rho_phi);                 // here: the INCREMENT Phi(x) - x
            y[n + i] = eval_inc(ct.v[i], p, hs) + sym(mj.rho_phi);
            for (int cc = 0; cc < D; cc++) {
                Kc[i][cc] = eval_inc(ps.Pu[i][cc], p, hs) + sym(mj.rho_psi);   // DPhi(x) - I
                Kc[n + i][cc] = eval_inc(ps.Pv[i][cc], p, hs) + sym(mj.rho_psi);
                Jc[i][cc] = ival(cc == i ? 1.0 : 0.0) + Kc[i][cc];
                Jc[n + i][cc] = ival(cc == n + i ? 1.0 : 0.0) + Kc[n + i][cc];
            }
        }
        // the mean-value form of the phi-set update needs DPhi on the segments [x, phi]: Jc + s dJ, s in [0,1]
        for (int i = 0; i < D; i++) for (int k = 0; k < D; k++) J[i][k] = Jc[i][k] + hull(dJ[i][k], ival(0.0));
        // phi' = (v, F(u)) over a sub-interval [tl,th] of the step and over the whole set (last column of the first/last frame,
        // topFrame.cpp:265-331):  phi(tau) in phi_c(tau) + Psi(tau; xi)(phi0 - x),  |Psi(tau;xi) - Psi_c(tau)| <= g1 tau |phi0 - x|_1
        auto phi_prime = [&](double tl, double th, ival* du, ival* dv) {
            ival dev[MAXD]; double dn = 0;
            for (int i = 0; i < D; i++) { ival a(0.0); for (int j = 0; j < D; j++) a = a + ival(C[i][j]) * r0[j]; dev[i] = a + r[i]; dn = add_ru(dn, mag(dev[i])); }
            const double pb = add_ru(mj.rho_psi, mul_ru(mul_ru(m2.g1, hs), dn));
            ival uu[MAXN], gg[MAXN];
            for (int i = 0; i < n; i++) {
                ival a = eval_rg(ct.u[i], p, tl, th) + sym(mj.rho_phi), bq = eval_rg(ct.v[i], p, tl, th) + sym(mj.rho_phi);
                for (int cc = 0; cc < D; cc++) {
                    a = a + (eval_rg(ps.Pu[i][cc], p, tl, th) + sym(pb)) * dev[cc];
                    bq = bq + (eval_rg(ps.Pv[i][cc], p, tl, th) + sym(pb)) * dev[cc];
                }
                uu[i] = a; du[i] = bq; gg[i] = a * (ival(1.0) - a);
            }
            for (int i = 0; i < n; i++) {
                const ival wi = uu[i] - ival(0.5);
                ival f = -((gg[i] * wi) * b[i]);
                if (i > 0)     f = f - (gg[i - 1] * wi) * c[i - 1];
                if (i < n - 1) f = f - (gg[i + 1] * wi) * c[i];
                dv[i] = f;
            }
        };
        if (s == 0) {     // topFrame::getFirstFrame (topFrame.cpp:301-331)
            for (int rr = 0; rr < D; rr++) for (int k = 0; k < n; k++) res.first_frame[rr * (CCP_MAX_N + 1) + k] = toc(Ws[rr][k]);
            ival du[MAXN], dv[MAXN]; if (g > 0) phi_prime(gp[0], gp[1], du, dv);
            if (g > 0) for (int i = 0; i < n; i++) {
                res.first_frame[i * (CCP_MAX_N + 1) + n] = toc(du[i]);
                res.first_frame[(n + i) * (CCP_MAX_N + 1) + n] = toc(dv[i]);
            }
        }
        if (last && g > 0) {
            const int ld = CCP_MAX_N + 2;
            ival du[MAXN], dv[MAXN]; phi_prime(gp[g - 1], gp[g], du, dv);
            for (int i = 0; i < n; i++) {
                res.last_frame[i * ld + n] = toc(du[i]);
                res.last_frame[(n + i) * ld + n] = toc(dv[i]);
            }
        }
        if (res.status != CCP_OK) break;          // non-finite determinant: stop before the set update
        // ---- set update of the front:  C+ = mid(J C),  r+ = J r + (y - x+) + (J C - C+) r0,  J = Jc + [0,1] dJ;  dot products in (mid,t) form ----
        double xn[MAXD], Cn[MAXD][MAXD]; ival rn[MAXD]; mt JCr[MAXD][MAXD], Jm[MAXD][MAXD], Jcm[MAXD][MAXD];
        for (int i = 0; i < D; i++) for (int k = 0; k < D; k++) { Jm[i][k] = iv2mt(J[i][k]); Jcm[i][k] = iv2mt(Jc[i][k]); }
        ival zc[MAXD];
        for (int i = 0; i < D; i++) { xn[i] = add_ru(x[i], mid_ru(y[i])); zc[i] = (ival(x[i]) - ival(xn[i])) + y[i]; }
        for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) {
            ival jc(0.0); for (int k = 0; k < D; k++) ipfma(jc, J[i][k], C[k][j]);
            Cn[i][j] = mid_ru(jc); JCr[i][j] = iv2mt(jc - ival(Cn[i][j]));
        }
        for (int i = 0; i < D; i++) {
            acc4 a; for (int j = 0; j < D; j++) acc_term(a, JCr[i][j], r0m[j]);
            for (int j = 0; j < D; j++) acc_term(a, Jm[i][j], iv2mt(r[j]));
            rn[i] = zc[i] + acc_finish(a);
        }
        // ---- frame update:  T+ = (Jc + sum_j Xi_j r0_j + Xi_r)(Tc + sum_l Theta_l r0_l + R) ----
        double Tcn[MAXD][MAXN]; ival Rn[MAXD][MAXN], Psn[MAXD][MAXN];
        static thread_local double Thn[MAXD][MAXD][MAXN];
        for (int i = 0; i < D; i++) for (int k = 0; k < n; k++) {
            ival jt(0.0); for (int l = 0; l < D; l++) ipfma(jt, Kc[i][l], Tc[l][k]);               // increment Kc Tc
            Tcn[i][k] = add_ru(Tc[i][k], mid_ru(jt));
            jt = (ival(Tc[i][k]) - ival(Tcn[i][k])) + jt;                                           // Jc Tc - Tc+
            acc4 e;
            for (int l = 0; l < D; l++) acc_term(e, Jcm[i][l], iv2mt(R[l][k]));                    // Jc R
            for (int l = 0; l < D; l++) acc_term(e, dJm[i][l], iv2mt(S0[l][k] + R[l][k]));         // second order
            ival xr(0.0); for (int l = 0; l < D; l++) ipfma(xr, Xi[D][i][l], Tc[l][k]);            // Xi_r Tc
            Rn[i][k] = (jt + acc_finish(e)) + xr;
        }
        for (int j = 0; j < D; j++) for (int i = 0; i < D; i++) for (int k = 0; k < n; k++) {
            ival t(0.0);
            for (int l = 0; l < D; l++) ipfma(t, Jc[i][l], Th[j][l][k]);
            for (int l = 0; l < D; l++) ipfma(t, Xi[j][i][l], Tc[l][k]);
            Thn[j][i][k] = mid_ru(t);
            Rn[i][k] = Rn[i][k] + (t - ival(Thn[j][i][k])) * r0[j];
        }
        for (int i = 0; i < D; i++) for (int k = 0; k < n; k++) {
            acc4 a; for (int l = 0; l < D; l++) acc_term(a, Jm[i][l], iv2mt(Ps[l][k]));
            Psn[i][k] = acc_finish(a);
        }
        for (int i = 0; i < D; i++) {
            x[i] = xn[i]; r[i] = rn[i];
            for (int j = 0; j < D; j++) C[i][j] = Cn[i][j];
            for (int k = 0; k < n; k++) { Tc[i][k] = Tcn[i][k]; R[i][k] = Rn[i][k]; Ps[i][k] = Psn[i][k]; for (int j = 0; j < D; j++) Th[j][i][k] = Thn[j][i][k]; }
        }
        tprev = ival(add_rd(tprev.lo, hs), add_ru(tprev.hi, hs));
        if (g > 0 && !last && ((s + 1) & renorm_mask) == 0) {
            double Hm[MAXN][MAXN];
            if (renorm_matrix(n, Tc, Hm)) {
                double Tc2[MAXD][MAXN]; ival R2[MAXD][MAXN];
                static thread_local double Th2[MAXD][MAXD][MAXN];
                for (int i = 0; i < D; i++) for (int k = 0; k < n; k++) {
                    ival t(0.0); for (int l = 0; l <= k; l++) ipfma(t, ival(Tc[i][l]), Hm[l][k]);
                    Tc2[i][k] = mid_ru(t);
                    ival rr = t - ival(Tc2[i][k]);
                    for (int l = 0; l <= k; l++) ipfma(rr, R[i][l], Hm[l][k]);
                    for (int j = 0; j < D; j++) {
                        ival u(0.0); for (int l = 0; l <= k; l++) ipfma(u, ival(Th[j][i][l]), Hm[l][k]);
                        Th2[j][i][k] = mid_ru(u);
                        rr = rr + (u - ival(Th2[j][i][k])) * r0[j];
                    }
                    R2[i][k] = rr;
                }
                ival Err2[MAXN][MAXN], Hi[MAXN][MAXN], G2[MAXN][MAXN];
                for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) { ival t(0.0); for (int l = 0; l <= k; l++) ipfma(t, Err[j][l], Hm[l][k]); Err2[j][k] = t; }
                for (int k = 0; k < n; k++) {                     // interval inverse of the upper triangular point matrix H
                    for (int i = 0; i < n; i++) Hi[i][k] = ival(0.0);
                    Hi[k][k] = ival(div_rd(1.0, Hm[k][k]), div_ru(1.0, Hm[k][k]));
                    for (int i = k - 1; i >= 0; i--) {
                        ival a(0.0); for (int m = i + 1; m <= k; m++) ipfma(a, Hi[m][k], Hm[i][m]);
                        Hi[i][k] = (-a) / ival(Hm[i][i]);
                    }
                }
                for (int i = 0; i < n; i++) for (int k = 0; k < n; k++) { ival a(0.0); for (int l = i; l < n; l++) a = a + Hi[i][l] * G[l][k]; G2[i][k] = a; }
                for (int k = 0; k < n; k++) sc = sc * Hi[k][k];
                for (int i = 0; i < D; i++) for (int k = 0; k < n; k++) { Tc[i][k] = Tc2[i][k]; R[i][k] = R2[i][k]; for (int j = 0; j < D; j++) Th[j][i][k] = Th2[j][i][k]; }
                for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) { Err[j][k] = Err2[j][k]; Errm[j][k] = iv2mt(Err2[j][k]); G[j][k] = G2[j][k]; }
                renorms++;
            }
        }
        if (last) {   // topFrame::getLastFrame (topFrame.cpp:265-299): frame and front point at the end of the last step
            const int ld = CCP_MAX_N + 2;
            ival S1[MAXD][MAXN], Wl[MAXD][MAXN];
            frame_of_state(n, Tc, Th, R, Ps, r0, Errm, S1, Wl);
            for (int rr = 0; rr < D; rr++) for (int k = 0; k < n; k++) {
                ival wl = Wl[rr][k], tu = (ival(Tc[rr][k]) + S1[rr][k]) + R[rr][k];
                if (renorms) {
                    // back to the basis A_u of the descriptor, piece by piece so that the affine dependence on r0 survives:
                    //   T = (Tc G) + sum_j (Theta_j G) r0_j + R G ;   W = T + Ps Err_0   (Err_0 = the descriptor's Eu_err; Err = Err_0 G^-1)
                    ival tc(0.0), rg(0.0), sj(0.0);
                    for (int l = 0; l <= k; l++) ipfma(tc, G[l][k], Tc[rr][l]);
                    for (int j = 0; j < D; j++) {
                        ival th(0.0); for (int l = 0; l <= k; l++) ipfma(th, G[l][k], Th[j][rr][l]);
                        sj = sj + th * r0[j];
                    }
                    for (int l = 0; l <= k; l++) rg = rg + R[rr][l] * G[l][k];
                    tu = (tc + sj) + rg;
                    acc4 e; for (int j = 0; j < n; j++) acc_term(e, iv2mt(Ps[rr][j]), iv2mt(fromc(d.Eu_err[j * CCP_MAX_N + k])));
                    wl = tu + acc_finish(e);
                }
                res.last_frame[rr * ld + k] = toc(wl);
                res.flow_P[rr * CCP_MAX_DIM + k] = toc(tu);                                             // [D Phi_T(set)] A_u: unstable columns
                res.flow_P[rr * CCP_MAX_DIM + n + k] = toc(Ps[rr][k]);                                  // stable columns
            }
            for (int i = 0; i < D; i++) {
                ival a(x[i]);
                for (int j = 0; j < D; j++) ipfma(a, r0[j], C[i][j]);
                res.last_frame[i * ld + n + 1] = toc(a + r[i]);
                if (g == 0) res.last_frame[i * ld + n] = toc(ival(x[i]) + r[i]);      // flow mode: image of the centre of the set (rho = 0)
            }
        }
    }
    res.zero_count = cnt.zero_count; res.failure_count = cnt.failure_count; res.first_failure = cnt.first_failure;
    res.n_conj = cnt.n_conj;
    for (int i = 0; i < cnt.n_conj; i++) { res.conj_time[i] = cnt.conj_time[i]; res.conj_index[i] = cnt.conj_index[i]; }
}

} // namespace c2
} // namespace orc
