// oracle/c1.hpp -- TEST INFRASTRUCTURE (CPU oracle, mode "c1").  Not part of the product path.
//
// Serial, plainly written restatement of the *structured* algorithm the CUDA kernel implements
// (DESIGN.md section 3).  It exists so that the GPU result can be compared BIT FOR BIT with a CPU
// computation performing the same directed-rounding operations in the same order (SURVEY.md
// Appendix C), while mode "ref" (oracle/ref.hpp) restates the reference's own structure.
//
// Mathematics: with phi=(u,v), W=(U,V) the reference's d=4n system
// (f_linearize, ODE_functions.cpp:206,239) is  phi' = f(phi),  W' = Df(phi) W,  so every
// eigen-direction k satisfies W_k(t) = Dphi_t(phi_0) W_k(0): the frame is the flow derivative
// applied to the initial frame.  One C^1 (Lohner) integration of the 2n-dim front ODE therefore
// replaces the reference's n separate 4n-dim C^0 integrations (propagateManifold.cpp:153-156):
//   * phi-set:   doubleton  x + C r0 + r   (C r0 = exact linear image of getPointNbd's box)
//   * P(t) = Dphi_t * A_u  (2n x 2n interval matrix), P+ = J P with J the step Jacobian over the hull
//   * frame:     W(t) = P(t) [I; Err]   (Err = Eu_m_Error_Final box, propagateManifold.cpp:48-55)
// Per step, Taylor tables (order p) of phi at the centre and over the hull, and of the variational
// matrix Psi (from the identity) over the hull; frame tables F_k = Psi^u_k W_s; then the reference's
// grid loop (utils.cpp:219-252) on F, the dets and the countZeros decision tree.
//
// Number formats: Cauchy products run in "mt" form (mid m, t >= |m|+rad) with 4 directed FMAs per
// term; everything else is inf-sup.  Truncation: a-priori box G + scalar majorant series, computed
// once per front (see c1_majorant).
#pragma once
#include "common.hpp"

namespace orc {
namespace c1 {

constexpr int MAXN = CCP_MAX_N, MAXD = CCP_MAX_DIM, MAXP = CCP_MAX_ORDER, MAXG = CCP_MAX_GRID;

struct mt { double m, t; };

static inline ival mt2iv(const mt& a) { double r = sub_ru(a.t, std::fabs(a.m)); return ival(sub_rd(a.m, r), add_ru(a.m, r)); }
static inline double mid_ru(const ival& a) { return add_ru(0.5 * a.lo, 0.5 * a.hi); }
// m = RU(lo/2 + hi/2) is never below the exact midpoint, hence m - lo >= hi - m and r = RU(m - lo) bounds both distances
// (bit-identical to max(RU(hi - m), RU(m - lo)) for normal numbers; see csrc/ccp_interval.cuh)
static inline mt iv2mt(const ival& a) {
    mt o; o.m = fma_ru(0.5, a.lo, mul_ru(0.5, a.hi));            // = RU(lo/2 + hi/2) with lo/2 unrounded: the kernel's instruction sequence
    const double r = sub_ru(o.m, a.lo);
    o.t = add_ru(std::fabs(o.m), r);
    return o;
}
struct acc4 { double slo = 0, shi = 0, T = 0, Q = 0; };
static inline void acc_term(acc4& a, const mt& x, const mt& y) {
    a.slo = fma_rd(x.m, y.m, a.slo);
    a.shi = fma_ru(x.m, y.m, a.shi);
    a.T   = fma_ru(x.t, y.t, a.T);
    a.Q   = fma_rd(std::fabs(x.m), std::fabs(y.m), a.Q);
}
// R = RU(T - Q) >= 0 always (termwise t_x t_y >= |m_x||m_y|, T rounded up, Q rounded down): no sign guard needed
static inline ival acc_finish(const acc4& a) {
    const double R = sub_ru(a.T, a.Q);
    return ival(sub_rd(a.slo, R), add_ru(a.shi, R));
}
// Summation order of every Cauchy product sum_j x[j] y[k-j]:  interior terms j = 1..k-1 ascending, then the two end
// terms j = 0 and j = k.  The end terms are the only ones that involve the order-k coefficients, so the GPU kernel can
// accumulate the interior of order k+1 in the same pass (same operands in registers) as the sum of order k.
template <class F> static inline void cauchy_order(int k, F term) {
    for (int j = 1; j < k; j++) term(j);
    term(0);
    if (k >= 1) term(k);
}
static inline void cauchy(acc4& a, const mt* x, const mt* y, int k) { cauchy_order(k, [&](int j) { acc_term(a, x[j], y[k - j]); }); }
// a / k for an integer k >= 1, as a product with the enclosure [rd(1/k), ru(1/k)] (no division in the hot loop:
// a directed FP64 division is a ~100-instruction dependent chain on the GPU).  One ulp wider than a true division.
struct InvK { double lo[MAXP + 3], hi[MAXP + 3]; InvK() { for (int k = 1; k < MAXP + 3; k++) { lo[k] = div_rd(1.0, (double)k); hi[k] = div_ru(1.0, (double)k); } } };
static inline ival divk(const ival& a, int k) {
    static thread_local InvK* inv = nullptr;
    if (!inv) inv = new InvK();       // built under FE_UPWARD by the calling oracle thread
    return ival(mul_rd(a.lo, a.lo >= 0 ? inv->lo[k] : inv->hi[k]), mul_ru(a.hi, a.hi >= 0 ? inv->hi[k] : inv->lo[k]));
}
// a / k directly in (mid,t) form, k >= 1: no interval conversion, no sign selects.  With ih = RU(1/k) and ihs = RU(ih (1 + 2^-50)):
//   m' = RN(m ih),  t' = RU(t ihs + 2^-1000)  >=  |m'| + (t - |m|) ih + |m| (ih - 1/k) + rounding of m'   (2^-1000 covers subnormal products)
static inline mt scalek(const mt& a, int k) {
    static thread_local InvK* inv = nullptr;
    if (!inv) inv = new InvK();
    mt o;
    { RoundNearestGuard rn; o.m = a.m * inv->hi[k]; }
    o.t = fma_ru(a.t, mul_ru(inv->hi[k], 1.0000000000000009), 0x1p-1000);
    return o;
}
static inline ival mulk(const ival& a, int k) { return ival(mul_rd((double)k, a.lo), mul_ru((double)k, a.hi)); }

// Horner steps for tau >= 0
static inline ival horner_pt(const ival& acc, double tau, const ival& c) { return ival(fma_rd(acc.lo, tau, c.lo), fma_ru(acc.hi, tau, c.hi)); }
static inline ival horner_rg(const ival& acc, double tl, double th, const ival& c) {
    return ival(fma_rd(acc.lo, acc.lo >= 0 ? tl : th, c.lo), fma_ru(acc.hi, acc.hi >= 0 ? th : tl, c.hi));
}

struct Majorant { double rho_phi, rho_phi_d, rho_psi, rho_psi_d; ival Gs_u, Gs_v; bool ok; };

// A-priori box G = {u in [-1/16,17/16], |v|<=1/2} and scalar majorant series for the Taylor
// coefficients of (phi,Psi) started anywhere in G (|Psi entries| <= beta_0).  All round-up.
static inline Majorant c1_majorant(int n, const ival* b, const ival* c, int p, double h) {
    Majorant M;
    double bh = 0, cs = 0;
    for (int i = 0; i < n; i++) bh = std::max(bh, mag(b[i]));
    for (int i = 0; i < n; i++) { double s = 0; if (i > 0) s = add_ru(s, mag(c[i - 1])); if (i < n - 1) s = add_ru(s, mag(c[i])); cs = std::max(cs, s); }
    const double W0 = 0.5625, G0 = 0.25, V0 = 0.5;
    double bc = add_ru(bh, cs);
    double F0 = mul_ru(mul_ru(bc, G0), W0);
    M.Gs_u = ival(add_ru(-0.0625, mul_ru(h, V0)), sub_rd(1.0625, mul_ru(h, V0)));
    M.Gs_v = ival(add_ru(-0.5, mul_ru(h, F0)), sub_rd(0.5, mul_ru(h, F0)));
    double w[MAXP + 3], g[MAXP + 3], be[MAXP + 3], Mh[MAXP + 3];
    w[0] = W0; g[0] = G0; w[1] = std::max(V0, F0);
    for (int k = 1; k <= p; k++) {
        double s = 0; for (int j = 0; j <= k; j++) s = fma_ru(w[j], w[k - j], s);
        g[k] = s;
        double f = 0; for (int j = 0; j <= k; j++) f = fma_ru(g[j], w[k - j], f);
        f = mul_ru(bc, f);
        w[k + 1] = div_ru(std::max(w[k], f), (double)(k + 1));
    }
    double hp = 1.0; for (int k = 0; k < p; k++) hp = mul_ru(hp, h);          // h^p
    M.rho_phi   = mul_ru(w[p + 1], mul_ru(hp, h));
    M.rho_phi_d = mul_ru(mul_ru((double)(p + 1), w[p + 1]), hp);
    double M0 = add_ru(add_ru(mul_ru(bh, 0.75), mul_ru(cs, 0.25)), mul_ru(mul_ru(2.0, cs), mul_ru(W0, W0)));
    double a = std::max(1.0, M0), ah = mul_ru(a, h);
    be[0] = add_ru(add_ru(1.0, ah), mul_ru(ah, ah));
    Mh[0] = M0;
    double k3 = mul_ru(3.0, bc);
    for (int k = 1; k <= p; k++) Mh[k] = mul_ru(k3, g[k]);
    for (int k = 0; k <= p; k++) {
        double s = 0; for (int j = 0; j <= k; j++) s = fma_ru(Mh[j], be[k - j], s);
        be[k + 1] = div_ru(std::max(be[k], s), (double)(k + 1));
    }
    M.rho_psi   = mul_ru(be[p + 1], mul_ru(hp, h));
    M.rho_psi_d = mul_ru(mul_ru((double)(p + 1), be[p + 1]), hp);
    M.ok = (ah <= 1.0) && std::isfinite(M.rho_phi) && std::isfinite(M.rho_psi) && M.Gs_u.lo < M.Gs_u.hi && M.Gs_v.lo < M.Gs_v.hi;
    return M;
}

struct PhiTables {               // one Taylor table set of the front ODE
    mt u[MAXN][MAXP + 1], v[MAXN][MAXP + 1], g[MAXN][MAXP + 1], q[MAXN][MAXP + 1], h[MAXN][MAXP + 1];
    mt w0[MAXN], z0[MAXN];       // order-0 terms of w = u - 1/2 and z = 1 - u (orders >= 1: w_k = u_k, z_k = -u_k)
};
static inline void acc_term_neg(acc4& a, const mt& x, const mt& y) {   // += x * (-y)
    a.slo = fma_rd(x.m, -y.m, a.slo);
    a.shi = fma_ru(x.m, -y.m, a.shi);
    a.T   = fma_ru(x.t, y.t, a.T);
    a.Q   = fma_rd(std::fabs(x.m), std::fabs(y.m), a.Q);
}

// Taylor coefficients of (u,v) up to order p from order-0 values (u0,v0).
// g = u(1-u) is formed as the series product u*z (NOT 1/4 - w^2): it keeps full RELATIVE precision where
// the front sits for most of [-L_-,0] (u ~ 1e-6 near the equilibrium 0), exactly like the reference's
// formula u*(u-.5)*(1-u) (ODE_functions.cpp:239); errors made there are amplified by e^{0.7 t}.
static inline void phi_tables(PhiTables& t, int n, int p, const ival* b, const ival* c, const ival* u0, const ival* v0) {
    for (int i = 0; i < n; i++) {
        t.u[i][0] = iv2mt(u0[i]); t.v[i][0] = iv2mt(v0[i]);
        t.w0[i] = iv2mt(u0[i] - ival(0.5)); t.z0[i] = iv2mt(ival(1.0) - u0[i]);
    }
    for (int k = 0; k < p; k++) {
        for (int i = 0; i < n; i++) {                 // g_i[k] = (u*z)[k],  z[m] = -u[m] for m >= 1
            acc4 a;
            cauchy_order(k, [&](int j) { if (k - j == 0) acc_term(a, t.u[i][j], t.z0[i]); else acc_term_neg(a, t.u[i][j], t.u[i][k - j]); });
            t.g[i][k] = iv2mt(acc_finish(a));
        }
        for (int e = 0; e + 1 < n; e++) {             // q_e[k] = (w_e*w_{e+1})[k]
            acc4 a;
            cauchy_order(k, [&](int j) { acc_term(a, j == 0 ? t.w0[e] : t.u[e][j], k - j == 0 ? t.w0[e + 1] : t.u[e + 1][k - j]); });
            t.q[e][k] = iv2mt(acc_finish(a));
        }
        for (int i = 0; i < n; i++) {                 // H_i[k] = -(b_i g_i + c_{i-1} g_{i-1} + c_i g_{i+1})[k]:  v' = F(u) = H w  (one product per component)
            const int im = i > 0 ? i - 1 : 0, ip = i < n - 1 ? i + 1 : n - 1;      // missing neighbours: zero parameter, clamped index
            const mt z = {0.0, 0.0};
            acc4 a;                                   // one (mid,t) dot product of parameters and g (the GPU's M/H lanes)
            acc_term(a, iv2mt(b[i]), t.g[i][k]); acc_term(a, i > 0 ? iv2mt(c[i - 1]) : z, t.g[im][k]); acc_term(a, i < n - 1 ? iv2mt(c[i]) : z, t.g[ip][k]);
            t.h[i][k] = iv2mt(-acc_finish(a));
        }
        for (int i = 0; i < n; i++) {
            acc4 a;
            cauchy_order(k, [&](int j) { acc_term(a, t.h[i][j], k - j == 0 ? t.w0[i] : t.u[i][k - j]); });
            t.v[i][k + 1] = scalek(iv2mt(acc_finish(a)), k + 1);
            t.u[i][k + 1] = k == 0 ? t.v[i][0] : scalek(t.v[i][k], k + 1);              // u[1] = v[0] exactly
        }
    }
}

struct PsiTables {
    mt Md[MAXN][MAXP + 1], Mo[MAXN][MAXP + 1];
    mt Pu[MAXN][MAXD][MAXP + 1], Pv[MAXN][MAXD][MAXP + 1];    // [row i][column c][order]
};

static inline void psi_tables(PsiTables& s, const PhiTables& hu, int n, int p, const ival* b, const ival* c) {
    const int D = 2 * n;
    for (int i = 0; i < n; i++) for (int cc = 0; cc < D; cc++) {
        s.Pu[i][cc][0] = iv2mt(ival(cc == i ? 1.0 : 0.0));
        s.Pv[i][cc][0] = iv2mt(ival(cc == n + i ? 1.0 : 0.0));
    }
    for (int k = 0; k < p; k++) {
        for (int i = 0; i < n; i++) {                 // Md_i[k] = [k = 0] b_i/2 - (3 b_i g_i + c_{i-1} g_{i-1} + c_i g_{i+1})[k],  Mo_e[k] = 2 c_e q_e[k]
            const int im = i > 0 ? i - 1 : 0, ip = i < n - 1 ? i + 1 : n - 1;
            const mt z = {0.0, 0.0};
            acc4 a;
            acc_term(a, iv2mt(mulk(b[i], 3)), hu.g[i][k]); acc_term(a, i > 0 ? iv2mt(c[i - 1]) : z, hu.g[im][k]); acc_term(a, i < n - 1 ? iv2mt(c[i]) : z, hu.g[ip][k]);
            ival d = -acc_finish(a);
            if (k == 0) d = ival(0.5 * b[i].lo, 0.5 * b[i].hi) + d;
            s.Md[i][k] = iv2mt(d);
            if (i < n - 1) {
                acc4 e;
                acc_term(e, iv2mt(mulk(c[i], 2)), hu.q[i][k]); acc_term(e, z, hu.q[i][k]); acc_term(e, z, hu.q[i][k]);
                s.Mo[i][k] = iv2mt(acc_finish(e));
            }
        }
        for (int i = 0; i < n; i++) for (int cc = 0; cc < D; cc++) {
            acc4 a0; cauchy(a0, s.Md[i], s.Pu[i][cc], k);
            ival sm = acc_finish(a0);                     // each product is enclosed on its own (one GPU lane each), then summed
            if (i > 0)     { acc4 a1; cauchy(a1, s.Mo[i - 1], s.Pu[i - 1][cc], k); sm = sm + acc_finish(a1); }
            if (i < n - 1) { acc4 a2; cauchy(a2, s.Mo[i], s.Pu[i + 1][cc], k); sm = sm + acc_finish(a2); }
            s.Pv[i][cc][k + 1] = scalek(iv2mt(sm), k + 1);
            s.Pu[i][cc][k + 1] = k == 0 ? s.Pv[i][cc][0] : scalek(s.Pv[i][cc][k], k + 1);
        }
    }
}

static inline ival eval_pt(const mt* s, int p, double tau) {
    ival acc = mt2iv(s[p]);
    for (int k = p - 1; k >= 0; k--) acc = horner_pt(acc, tau, mt2iv(s[k]));
    return acc;
}
static inline ival eval_rg(const mt* s, int p, double tl, double th) {
    ival acc = mt2iv(s[p]);
    for (int k = p - 1; k >= 0; k--) acc = horner_rg(acc, tl, th, mt2iv(s[k]));
    return acc;
}
static inline ival eval_rg_deriv(const mt* s, int p, double tl, double th) {
    ival acc = mulk(mt2iv(s[p]), p);
    for (int k = p - 1; k >= 1; k--) acc = horner_rg(acc, tl, th, mulk(mt2iv(s[k]), k));
    return acc;
}

struct Output {
    ccp_front_result res;
    std::vector<ccp_series_rec>* series = nullptr;
};

static inline ccp_interval toc(const ival& a) { ccp_interval o; o.lo = a.lo; o.hi = a.hi; return o; }
static inline ival fromc(const ccp_interval& a) { return ival(a.lo, a.hi); }

static inline void run_front(const ccp_front_desc& d, ccp_front_result& res, std::vector<ccp_series_rec>* series) {
    std::memset(&res, 0, sizeof(res));
    res.first_failure = -1;
    const int n = d.n, D = 2 * n, p = d.order, g = d.grid;
    if (n < 2 || n > MAXN || p < 2 || p > MAXP || g < 1 || g > MAXG || d.step_log2 < 1 || d.step_log2 > 20 || !(d.L_minus > 0) || !std::isfinite(d.L_minus) || std::ldexp(d.L_minus, d.step_log2) > 16777216.0) {
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

    // ---- initial sets (construct_InitCondU / construct_A_lin, propagateManifold.cpp:6-81) ----
    double x[MAXD], C[MAXD][MAXD];
    ival r0[MAXD], r[MAXD], P[MAXD][MAXD], Err[MAXN][MAXN];
    for (int i = 0; i < D; i++) {
        ival pi = fromc(d.p_i[i]);
        x[i] = mid_ru(pi); r[i] = pi - ival(x[i]); r0[i] = fromc(d.Uxy[i]);
        for (int j = 0; j < D; j++) { C[i][j] = d.A_u[i * CCP_MAX_DIM + j]; P[i][j] = ival(C[i][j]); }
    }
    for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) Err[j][k] = fromc(d.Eu_err[j * CCP_MAX_N + k]);

    Counter cnt;
    ival tprev(0.0);
    static thread_local PhiTables ct, ht;
    static thread_local PsiTables ps;
    ival Fk[MAXP + 1][MAXN][MAXN];
    ival FRprev[MAXN][MAXN];

    for (int s = 0; s < S; s++) {
        const bool last = (s == S - 1);
        const double hs = last ? sub_ru(d.L_minus, tprev.hi) : h;       // exact (Sterbenz) for the reference's T, h
        // hull of the phi-set
        ival X[MAXD];
        for (int i = 0; i < D; i++) {
            ival a(x[i]);
            for (int j = 0; j < D; j++) a = a + ival(C[i][j]) * r0[j];
            X[i] = a + r[i];
        }
        bool inside = true, fin = true;
        for (int i = 0; i < n; i++) {
            inside = inside && subset(X[i], mj.Gs_u) && subset(X[n + i], mj.Gs_v);
            fin = fin && finite(X[i]) && finite(X[n + i]);
        }
        if (!fin) { res.status = CCP_ST_NONFINITE; break; }
        if (!inside) { res.status = CCP_ST_ENCLOSURE; break; }
        // tables
        ival xc[MAXD]; for (int i = 0; i < D; i++) xc[i] = ival(x[i]);
        phi_tables(ct, n, p, b, c, xc, xc + n);
        phi_tables(ht, n, p, b, c, X, X + n);
        psi_tables(ps, ht, n, p, b, c);
        // frame at step start and frame tables
        ival Ws[MAXD][MAXN]; mt Wm[MAXD][MAXN]; double rhoF[MAXN], rhoFd[MAXN];
        for (int cc = 0; cc < D; cc++) for (int bb = 0; bb < n; bb++) {
            ival a = P[cc][bb];
            for (int j = 0; j < n; j++) a = a + P[cc][n + j] * Err[j][bb];
            Ws[cc][bb] = a; Wm[cc][bb] = iv2mt(a);
        }
        for (int bb = 0; bb < n; bb++) {
            double sm = 0; for (int cc = 0; cc < D; cc++) sm = add_ru(sm, mag(Ws[cc][bb]));
            rhoF[bb] = mul_ru(mj.rho_psi, sm); rhoFd[bb] = mul_ru(mj.rho_psi_d, sm);
        }
        for (int k = 0; k <= p; k++) for (int a = 0; a < n; a++) for (int bb = 0; bb < n; bb++) {
            acc4 ac; for (int cc = 0; cc < D; cc++) acc_term(ac, ps.Pu[a][cc][k], Wm[cc][bb]);
            ival f = acc_finish(ac);
            if (k == 0) f = f + sym(rhoF[bb]);
            Fk[k][a][bb] = f;
        }
        // ---- grid loop (utils.cpp:219-252) + topFrame (topFrame.cpp:30-108,193-263) ----
        double gp[MAXG + 1]; grid_points(hs, g, gp);
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
                FD[a][bb] = acc + sym(rhoFd[bb]);
            }
            ival time(add_rd(tprev.lo, tl), add_ru(tprev.hi, th));
            ival dL = detN(FL, n), dR = detN(FR, n), dV = detN(FV, n);
            // the Jacobi derivative is formed lazily (as topFrame::countZeros does): only when the rough enclosures do not
            // settle the sub-interval, or when the series is requested
            const ival LR = dL * dR;
            const bool need_D = (series != nullptr) || !(LR.lo > 0 && definite(dV));
            ival dD = need_D ? jacobi_derivative(FV, FD, n) : ival(0.0);
            int idx = s * g + i;
            count_step(cnt, idx, time, dL, dR, dV, dD);
            if (series) { ccp_series_rec rec; rec.time = toc(time); rec.det_L = toc(dL); rec.det_R = toc(dR); rec.det_V = toc(dV); rec.det_D = toc(dD); series->push_back(rec); }
            if (idx == 0) res.det_first = toc(dV);
            if (idx == S * g - 1) res.det_last = toc(dV);
            if (!(finite(dL) && finite(dR) && finite(dV) && finite(dD))) res.status = CCP_ST_NONFINITE;
        }
        // first-frame data (topFrame::getFirstFrame, topFrame.cpp:301-331)
        if (s == 0) {
            for (int rr = 0; rr < D; rr++) for (int k = 0; k < n; k++) res.first_frame[rr * (CCP_MAX_N + 1) + k] = toc(Ws[rr][k]);
            for (int i = 0; i < n; i++) {
                ival du = eval_rg(ht.v[i], p, gp[0], gp[1]) + sym(mj.rho_phi);
                ival dv = eval_rg_deriv(ht.v[i], p, gp[0], gp[1]) + sym(mj.rho_phi_d);
                res.first_frame[i * (CCP_MAX_N + 1) + n] = toc(du);
                res.first_frame[(n + i) * (CCP_MAX_N + 1) + n] = toc(dv);
            }
        }
        // ---- step: y = Phi(x), J = DPhi over the hull ----
        ival y[MAXD], J[MAXD][MAXD];
        for (int i = 0; i < n; i++) {
            y[i] = eval_pt(ct.u[i], p, hs) + sym(mj.rho_phi);
            y[n + i] = eval_pt(ct.v[i], p, hs) + sym(mj.rho_phi);
            for (int cc = 0; cc < D; cc++) {
                J[i][cc] = eval_pt(ps.Pu[i][cc], p, hs) + sym(mj.rho_psi);
                J[n + i][cc] = eval_pt(ps.Pv[i][cc], p, hs) + sym(mj.rho_psi);
            }
        }
        if (last) {   // topFrame::getLastFrame (topFrame.cpp:265-299)
            const int ld = CCP_MAX_N + 2;
            for (int rr = 0; rr < D; rr++) for (int k = 0; k < n; k++) {
                ival a(0.0); for (int cc = 0; cc < D; cc++) a = a + J[rr][cc] * Ws[cc][k];
                res.last_frame[rr * ld + k] = toc(a);
            }
            for (int i = 0; i < n; i++) {
                ival du = eval_rg(ht.v[i], p, gp[g - 1], gp[g]) + sym(mj.rho_phi);
                ival dv = eval_rg_deriv(ht.v[i], p, gp[g - 1], gp[g]) + sym(mj.rho_phi_d);
                res.last_frame[i * ld + n] = toc(du);
                res.last_frame[(n + i) * ld + n] = toc(dv);
            }
        }
        if (res.status != CCP_OK) break;          // non-finite determinant: stop before the set update
        // set update
        double xn[MAXD], Cn[MAXD][MAXD]; ival rn[MAXD], JC[MAXD][MAXD], Pn[MAXD][MAXD];
        for (int i = 0; i < D; i++) xn[i] = mid_ru(y[i]);
        for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) {
            ival a(0.0); for (int k = 0; k < D; k++) a = a + J[i][k] * ival(C[k][j]);
            JC[i][j] = a; Cn[i][j] = mid_ru(a);
        }
        for (int i = 0; i < D; i++) {
            ival z = y[i] - ival(xn[i]);
            for (int j = 0; j < D; j++) z = z + (JC[i][j] - ival(Cn[i][j])) * r0[j];
            ival a(0.0); for (int j = 0; j < D; j++) a = a + J[i][j] * r[j];
            rn[i] = a + z;
        }
        for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) {
            ival a(0.0); for (int k = 0; k < D; k++) a = a + J[i][k] * P[k][j];
            Pn[i][j] = a;
        }
        for (int i = 0; i < D; i++) { x[i] = xn[i]; r[i] = rn[i]; for (int j = 0; j < D; j++) { C[i][j] = Cn[i][j]; P[i][j] = Pn[i][j]; } }
        tprev = ival(add_rd(tprev.lo, hs), add_ru(tprev.hi, hs));
        if (last) {
            const int ld = CCP_MAX_N + 2;
            for (int i = 0; i < D; i++) {
                ival a(x[i]);
                for (int j = 0; j < D; j++) a = a + ival(C[i][j]) * r0[j];
                res.last_frame[i * ld + n + 1] = toc(a + r[i]);
            }
        }
    }
    res.zero_count = cnt.zero_count; res.failure_count = cnt.failure_count; res.first_failure = cnt.first_failure;
    res.n_conj = cnt.n_conj;
    for (int i = 0; i < cnt.n_conj; i++) { res.conj_time[i] = cnt.conj_time[i]; res.conj_index[i] = cnt.conj_index[i]; }
}

} // namespace c1
} // namespace orc
