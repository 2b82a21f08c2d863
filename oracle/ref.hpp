// oracle/ref.hpp -- TEST INFRASTRUCTURE (CPU oracle, mode "ref").  Not part of the product path.
//
// Reference-structured restatement of the hot path:
//   propagateManifold::frameDet            propagateManifold.cpp:125-188
//   computeTotalTrajectory (x n)           propagateManifold.cpp:84-121
//   construct_InitCondU / construct_A_lin  propagateManifold.cpp:6-81
//   getTotalTrajectory + constructMatrix   utils.cpp:139-160,175-263
//   f_linearize                            ODE_functions.cpp:206 (n=2), :239 (n=3); n=4 = same chain pattern
//   topFrame series / dets / countZeros    topFrame.cpp:4-108,110-174,193-263
//   getLastFrame / getFirstFrame           topFrame.cpp:265-331
// The arithmetic the reference delegates to CAPD 5.0.59 (NOT vendored, not installable here) is
// restated from its published algorithms [CAPD-recollection, SURVEY.md 8c]:
//   IMap from a formula string  -> expression tape + automatic differentiation of Taylor series
//   ITaylor(order p), fixed step -> C^0 Lohner step on a doubleton x + C r0 + B r with QR-factorised B,
//                                   first-order rough enclosure, remainder a_{p+1}([E]) h^{p+1}
//   SolutionCurve(t), timeDerivative(t) -> Horner of centre/Jacobian tables in mean-value form
//   det -> cofactor expansion (common.hpp)
#pragma once
#include "common.hpp"
#include <thread>

namespace orc {
namespace ref {

constexpr int MAXN = CCP_MAX_N;

enum Op { VAR, CONSTANT, ADD, SUB, MUL, NEG, SCALE };
struct Node { Op op; int a, b; ival c; };
struct Tape {
    int d = 0;
    std::vector<Node> nodes;
    std::vector<int> rhs;
    std::vector<std::vector<int>> dirs;     // structurally non-zero gradient directions per node (sorted)
    int add(Op op, int a = -1, int b = -1, ival c = ival(0.0)) { nodes.push_back(Node{op, a, b, c}); return (int)nodes.size() - 1; }
};
struct E { Tape* t; int id; };
static inline E operator+(E a, E b) { return E{a.t, a.t->add(ADD, a.id, b.id)}; }
static inline E operator-(E a, E b) { return E{a.t, a.t->add(SUB, a.id, b.id)}; }
static inline E operator*(E a, E b) { return E{a.t, a.t->add(MUL, a.id, b.id)}; }
static inline E operator-(E a) { return E{a.t, a.t->add(NEG, a.id)}; }
static inline E scale(const ival& c, E a) { return E{a.t, a.t->add(SCALE, a.id, -1, c)}; }   // parameter or literal times expression
static inline E operator-(E a, double c) { return a - E{a.t, a.t->add(CONSTANT, -1, -1, ival(c))}; }
static inline E rsub(double c, E a) { return E{a.t, a.t->add(CONSTANT, -1, -1, ival(c))} - a; }

// f_linearize for the chain of n components, written term by term as in the formula strings.
static inline void build_tape(Tape& T, int n, const ival* b, const ival* c) {
    T = Tape(); T.d = 4 * n;
    for (int i = 0; i < 4 * n; i++) T.add(VAR);
    auto u = [&](int i) { return E{&T, i}; };
    auto v = [&](int i) { return E{&T, n + i}; };
    auto U = [&](int i) { return E{&T, 2 * n + i}; };
    auto V = [&](int i) { return E{&T, 3 * n + i}; };
    T.rhs.assign(4 * n, -1);
    for (int i = 0; i < n; i++) T.rhs[i] = v(i).id;
    for (int i = 0; i < n; i++) {
        // -b_i*u_i*(u_i-.5)*(1-u_i) - c_ij*u_j*(1-u_j)*(u_i-.5) ...
        E e = scale(-b[i], u(i)) * (u(i) - 0.5) * rsub(1.0, u(i));
        if (i > 0)     e = e - scale(c[i - 1], u(i - 1)) * rsub(1.0, u(i - 1)) * (u(i) - 0.5);
        if (i < n - 1) e = e - scale(c[i], u(i + 1)) * rsub(1.0, u(i + 1)) * (u(i) - 0.5);
        T.rhs[n + i] = e.id;
    }
    for (int i = 0; i < n; i++) T.rhs[2 * n + i] = V(i).id;
    for (int i = 0; i < n; i++) {
        // (-b_i*(3*u_i*(1-u_i)-1/2) - c_ij*u_j*(1-u_j))*U_i  -  c_ij*(-2*(u_i-1/2)*(u_j-1/2))*U_j
        E dg = scale(-b[i], (scale(ival(3.0), u(i)) * rsub(1.0, u(i))) - 0.5);
        if (i > 0)     dg = dg - scale(c[i - 1], u(i - 1)) * rsub(1.0, u(i - 1));
        if (i < n - 1) dg = dg - scale(c[i], u(i + 1)) * rsub(1.0, u(i + 1));
        E e = dg * U(i);
        bool have = false; E acc = e;
        if (i > 0) {
            E off = scale(-c[i - 1], scale(ival(-2.0), (u(i - 1) - 0.5)) * (u(i) - 0.5)) * U(i - 1);
            acc = off + e; have = true;
        }
        if (!have) acc = e;
        if (i < n - 1) {
            E off = scale(c[i], scale(ival(-2.0), (u(i) - 0.5)) * (u(i + 1) - 0.5)) * U(i + 1);
            acc = acc - off;
        }
        T.rhs[3 * n + i] = acc.id;
    }
    // structural sparsity of d(node)/d(initial condition): fixed point of the dependency closure
    const int N = (int)T.nodes.size(), d = T.d;
    std::vector<std::vector<char>> nz(N, std::vector<char>(d, 0));
    for (int i = 0; i < d; i++) nz[i][i] = 1;
    for (bool changed = true; changed;) {
        changed = false;
        for (int k = d; k < N; k++) {
            const Node& nd = T.nodes[k];
            for (int q = 0; q < d; q++) {
                char z = 0;
                if (nd.a >= 0) z |= nz[nd.a][q];
                if (nd.b >= 0) z |= nz[nd.b][q];
                if (z && !nz[k][q]) { nz[k][q] = 1; changed = true; }
            }
        }
        for (int i = 0; i < d; i++) for (int q = 0; q < d; q++) if (nz[T.rhs[i]][q] && !nz[i][q]) { nz[i][q] = 1; changed = true; }
    }
    T.dirs.assign(N, {});
    for (int k = 0; k < N; k++) for (int q = 0; q < d; q++) if (nz[k][q]) T.dirs[k].push_back(q);
}

// Taylor tables of every tape node: value series and (optionally) gradient series wrt the initial condition
struct Tables {
    int N = 0, d = 0, K = 0; bool grad = false;
    std::vector<ival> val;              // [node][K+1]
    std::vector<ival> g;                // [node][dir][K+1]  (dense in dir; zero where structurally zero)
    ival& V(int node, int k) { return val[(size_t)node * (K + 1) + k]; }
    ival& G(int node, int q, int k) { return g[((size_t)node * d + q) * (K + 1) + k]; }
};

static inline void compute_tables(const Tape& T, Tables& tb, const ival* init, int K, bool with_grad) {
    const int N = (int)T.nodes.size(), d = T.d;
    tb.N = N; tb.d = d; tb.K = K; tb.grad = with_grad;
    tb.val.assign((size_t)N * (K + 1), ival(0.0));
    if (with_grad) tb.g.assign((size_t)N * d * (K + 1), ival(0.0)); else tb.g.clear();
    for (int i = 0; i < d; i++) { tb.V(i, 0) = init[i]; if (with_grad) tb.G(i, i, 0) = ival(1.0); }
    for (int k = 0; k <= K; k++) {
        for (int id = d; id < N; id++) {
            const Node& nd = T.nodes[id];
            switch (nd.op) {
                case CONSTANT: if (k == 0) tb.V(id, 0) = nd.c; break;
                case ADD: tb.V(id, k) = tb.V(nd.a, k) + tb.V(nd.b, k);
                    if (with_grad) for (int q : T.dirs[id]) tb.G(id, q, k) = tb.G(nd.a, q, k) + tb.G(nd.b, q, k);
                    break;
                case SUB: tb.V(id, k) = tb.V(nd.a, k) - tb.V(nd.b, k);
                    if (with_grad) for (int q : T.dirs[id]) tb.G(id, q, k) = tb.G(nd.a, q, k) - tb.G(nd.b, q, k);
                    break;
                case NEG: tb.V(id, k) = -tb.V(nd.a, k);
                    if (with_grad) for (int q : T.dirs[id]) tb.G(id, q, k) = -tb.G(nd.a, q, k);
                    break;
                case SCALE: tb.V(id, k) = nd.c * tb.V(nd.a, k);
                    if (with_grad) for (int q : T.dirs[id]) tb.G(id, q, k) = nd.c * tb.G(nd.a, q, k);
                    break;
                case MUL: {
                    ival s(0.0);
                    for (int j = 0; j <= k; j++) s = s + tb.V(nd.a, j) * tb.V(nd.b, k - j);
                    tb.V(id, k) = s;
                    if (with_grad) {
                        for (int q : T.dirs[nd.a]) { ival t(0.0); for (int j = 0; j <= k; j++) t = t + tb.G(nd.a, q, j) * tb.V(nd.b, k - j); tb.G(id, q, k) = t; }
                        for (int q : T.dirs[nd.b]) { ival t = tb.G(id, q, k); for (int j = 0; j <= k; j++) t = t + tb.V(nd.a, j) * tb.G(nd.b, q, k - j); tb.G(id, q, k) = t; }
                    }
                    break;
                }
                default: break;
            }
        }
        if (k < K) for (int i = 0; i < d; i++) {
            ival kk((double)(k + 1));
            tb.V(i, k + 1) = tb.V(T.rhs[i], k) / kk;
            if (with_grad) for (int q : T.dirs[i]) tb.G(i, q, k + 1) = tb.G(T.rhs[i], q, k) / kk;
        }
    }
}

static inline void eval_field(const Tape& T, const ival* z, ival* out) {
    Tables tb; compute_tables(T, tb, z, 0, false);
    for (int i = 0; i < T.d; i++) out[i] = tb.V(T.rhs[i], 0);
}

// what constructMatrix (utils.cpp:139-160) packs, reduced to the rows topFrame reads every sub-interval
struct SubRec { ival time; ival U[MAXN][4]; };   // U[m][type]: type 0 left, 1 right, 2 range, 3 derivative
struct Trajectory {
    std::vector<SubRec> sub;
    std::vector<ival> first_full, last_full;      // d x 5 IMatrix of the first / last sub-interval (cols 1..4 used)
    int status = CCP_OK;
};

// Gram-Schmidt orthonormalisation of the columns of M (d x d), columns pre-sorted by size (Lohner).
static inline bool qr_orth(int d, const std::vector<double>& M, const std::vector<double>& weight, std::vector<double>& Q) {
    std::vector<int> order(d);
    std::vector<double> key(d);
    for (int j = 0; j < d; j++) { double s = 0; for (int i = 0; i < d; i++) s += M[i * d + j] * M[i * d + j]; key[j] = std::sqrt(s) * weight[j]; order[j] = j; }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] > key[b]; });
    Q.assign(d * d, 0.0);
    for (int jj = 0; jj < d; jj++) {
        int j = order[jj];
        std::vector<double> col(d);
        for (int i = 0; i < d; i++) col[i] = M[i * d + j];
        double n0 = 0; for (int i = 0; i < d; i++) n0 += col[i] * col[i];
        for (int pass = 0; pass < 2; pass++)
            for (int k = 0; k < jj; k++) {
                double dot = 0; for (int i = 0; i < d; i++) dot += Q[i * d + k] * col[i];
                for (int i = 0; i < d; i++) col[i] -= dot * Q[i * d + k];
            }
        double n1 = 0; for (int i = 0; i < d; i++) n1 += col[i] * col[i];
        if (!(n1 > 1e-24 * n0) || !(n1 > 0)) return false;
        n1 = std::sqrt(n1);
        for (int i = 0; i < d; i++) Q[i * d + jj] = col[i] / n1;
    }
    return true;
}
// rigorous enclosure of Q^-1 for a nearly orthogonal point matrix: Q^-1 = (I-E)^-1 Q^T, E = I - Q^T Q
static inline bool inverse_orth(int d, const std::vector<double>& Q, std::vector<ival>& Qi) {
    double eps = 0;
    for (int i = 0; i < d; i++) {
        double row = 0;
        for (int j = 0; j < d; j++) {
            ival s(0.0);
            for (int k = 0; k < d; k++) s = s + ival(Q[k * d + i]) * ival(Q[k * d + j]);
            ival e = ival(i == j ? 1.0 : 0.0) - s;
            row = add_ru(row, mag(e));
        }
        eps = std::max(eps, row);
    }
    if (!(eps < 0.5)) return false;
    double delta = div_ru(eps, sub_rd(1.0, eps));
    Qi.assign(d * d, ival(0.0));
    for (int j = 0; j < d; j++) {
        double cs = 0; for (int k = 0; k < d; k++) cs = add_ru(cs, std::fabs(Q[j * d + k]));   // sum_k |Q^T[k][j]| = sum_k |Q[j][k]|
        double e = mul_ru(delta, cs);
        for (int i = 0; i < d; i++) Qi[i * d + j] = ival(Q[j * d + i]) + sym(e);
    }
    return true;
}

static inline ival horner(const ival* coef, int stride, int p, const ival& tau) {
    ival acc = coef[(size_t)p * stride];
    for (int k = p - 1; k >= 0; k--) acc = acc * tau + coef[(size_t)k * stride];
    return acc;
}
static inline ival horner_deriv(const ival* coef, int stride, int p, const ival& tau) {
    ival acc = ival((double)p) * coef[(size_t)p * stride];
    for (int k = p - 1; k >= 1; k--) acc = acc * tau + ival((double)k) * coef[(size_t)k * stride];
    return acc;
}
static inline ival ipow(const ival& a, int e) { ival r(1.0); for (int i = 0; i < e; i++) r = r * a; return r; }

// computeTotalTrajectory(eigenvector_NUM = kdir, T = L_minus, grid)  (propagateManifold.cpp:84-121)
static inline void total_trajectory(const ccp_front_desc& D, const Tape& T, int kdir, Trajectory& out) {
    const int n = D.n, dim = 2 * n, d = 4 * n, p = D.order, g = D.grid;
    const double h = std::ldexp(1.0, -D.step_log2);
    // construct_InitCondU
    std::vector<ival> pt(d), nbd(d);
    for (int i = 0; i < dim; i++) { pt[i] = ival(D.p_i[i].lo, D.p_i[i].hi); nbd[i] = ival(D.Uxy[i].lo, D.Uxy[i].hi); }
    for (int i = 0; i < dim; i++) pt[dim + i] = ival(D.A_u[i * CCP_MAX_DIM + kdir]);
    for (int i = 0; i < n; i++) { nbd[dim + i] = ival(0.0); nbd[dim + n + i] = ival(D.Eu_err[i * CCP_MAX_N + kdir].lo, D.Eu_err[i * CCP_MAX_N + kdir].hi); }
    // construct_A_lin: blockdiag(A_u, A_u); C0Rect2Set(pt, A_lin, nbd) = mid(pt) + A_lin*nbd + I*(pt-mid(pt))
    std::vector<double> x(d), C(d * d, 0.0), B(d * d, 0.0);
    std::vector<ival> r0 = nbd, r(d), Binv(d * d, ival(0.0));
    for (int i = 0; i < dim; i++) for (int j = 0; j < dim; j++) { C[i * d + j] = D.A_u[i * CCP_MAX_DIM + j]; C[(dim + i) * d + dim + j] = D.A_u[i * CCP_MAX_DIM + j]; }
    for (int i = 0; i < d; i++) { x[i] = mid(pt[i]); r[i] = pt[i] - ival(x[i]); B[i * d + i] = 1.0; Binv[i * d + i] = ival(1.0); }

    const int S = step_count(D.L_minus, h);
    out.sub.clear(); out.sub.reserve((size_t)S * g);
    Tables ct, ht, et;
    ival prevTime(0.0);
    std::vector<ival> X(d), Delta(d), Eb(d), Y(d), F(d), y(d), rem(d), remd(d), J(d * d), JC(d * d), JB(d * d), Z(d), rn(d);
    for (int s = 0; s < S; s++) {
        const bool last = (s == S - 1);
        const ival stepMade = last ? ival(D.L_minus) - prevTime : ival(h);
        // hull and deviation from the centre
        for (int i = 0; i < d; i++) {
            ival a(0.0);
            for (int j = 0; j < d; j++) if (C[i * d + j] != 0.0) a = a + ival(C[i * d + j]) * r0[j];
            for (int j = 0; j < d; j++) if (B[i * d + j] != 0.0) a = a + ival(B[i * d + j]) * r[j];
            Delta[i] = a; X[i] = ival(x[i]) + a;
            if (!finite(X[i])) { out.status = CCP_ST_NONFINITE; return; }
        }
        // first-order rough enclosure  X + [0,h] f(E) subset E
        const ival zh = ival(0.0, stepMade.hi);
        eval_field(T, X.data(), F.data());
        for (int i = 0; i < d; i++) Eb[i] = X[i] + zh * F[i];
        bool ok = false;
        for (int it = 0; it < 40 && !ok; it++) {
            std::vector<ival> Ei(d);
            for (int i = 0; i < d; i++) { double e = add_ru(mul_ru(0.125, width(Eb[i])), mul_ru(1e-14, mag(Eb[i]))); e = add_ru(e, 1e-300); Ei[i] = ival(sub_rd(Eb[i].lo, e), add_ru(Eb[i].hi, e)); }
            eval_field(T, Ei.data(), F.data());
            ok = true;
            for (int i = 0; i < d; i++) { Y[i] = X[i] + zh * F[i]; if (!subset(Y[i], Ei[i])) ok = false; }
            if (ok) { for (int i = 0; i < d; i++) Eb[i] = Y[i]; }
            else    { for (int i = 0; i < d; i++) Eb[i] = hull(Eb[i], Y[i]); }
        }
        if (!ok) { out.status = CCP_ST_ENCLOSURE; return; }
        // Taylor tables: centre (values), hull (values+Jacobian), enclosure (values, order p+1)
        std::vector<ival> xc(d); for (int i = 0; i < d; i++) xc[i] = ival(x[i]);
        compute_tables(T, ct, xc.data(), p, false);
        compute_tables(T, ht, X.data(), p, true);
        compute_tables(T, et, Eb.data(), p + 1, false);

        auto curve = [&](const ival& tau, bool deriv, ival* res) {
            // SolutionCurve::operator()(t) / timeDerivative(t): centre polynomial + Jacobian polynomial * (set - centre) + remainder
            ival tp = deriv ? ival((double)(p + 1)) * ipow(tau, p) : ipow(tau, p + 1);
            for (int i = 0; i < d; i++) {
                ival a = deriv ? horner_deriv(&ct.V(i, 0), 1, p, tau) : horner(&ct.V(i, 0), 1, p, tau);
                for (int q : T.dirs[i]) {
                    ival jq = deriv ? horner_deriv(&ht.G(i, q, 0), 1, p, tau) : horner(&ht.G(i, q, 0), 1, p, tau);
                    a = a + jq * Delta[q];
                }
                res[i] = a + et.V(i, p + 1) * tp;
            }
        };
        // grid loop (utils.cpp:219-252)
        double gp[CCP_MAX_GRID + 1];
        {   // gridPoints[i] = (i*stepMade/grid).left(); gridPoints[grid] = stepMade
            for (int i = 0; i < g; i++) gp[i] = (ival((double)i) * stepMade / ival((double)g)).lo;
        }
        std::vector<ival> vl(d), vr(d), vv(d), vd(d);
        for (int i = 0; i < g; i++) {
            ival subd = hull(ival(gp[i]), (i + 1 < g) ? ival(gp[i + 1]) : stepMade);
            ival dom = ival(0.0, 1.0) * stepMade;
            if (!intersect(dom, subd, subd)) { out.status = CCP_ST_BAD_DESC; return; }
            curve(ival(subd.lo), false, vl.data());
            curve(ival(subd.hi), false, vr.data());
            curve(subd, false, vv.data());
            curve(subd, true, vd.data());
            SubRec rec; rec.time = prevTime + subd;
            for (int m = 0; m < n; m++) { rec.U[m][0] = vl[dim + m]; rec.U[m][1] = vr[dim + m]; rec.U[m][2] = vv[dim + m]; rec.U[m][3] = vd[dim + m]; }
            out.sub.push_back(rec);
            const bool is_first = (s == 0 && i == 0), is_last = (last && i == g - 1);
            if (is_first || is_last) {
                std::vector<ival>& dst = is_first ? out.first_full : out.last_full;
                dst.assign((size_t)d * 5, ival(0.0));
                dst[0] = rec.time;
                for (int rr = 0; rr < d; rr++) { dst[rr * 5 + 1] = vl[rr]; dst[rr * 5 + 2] = vr[rr]; dst[rr * 5 + 3] = vv[rr]; dst[rr * 5 + 4] = vd[rr]; }
                if (is_first && is_last) out.last_full = out.first_full;
            }
        }
        // ---- the C^0 Lohner step ----
        const ival hh = stepMade;
        const ival hp1 = ipow(hh, p + 1);
        for (int i = 0; i < d; i++) {
            y[i] = horner(&ct.V(i, 0), 1, p, hh) + et.V(i, p + 1) * hp1;
            for (int q = 0; q < d; q++) J[i * d + q] = ival(0.0);
            for (int q : T.dirs[i]) J[i * d + q] = horner(&ht.G(i, q, 0), 1, p, hh);
        }
        std::vector<double> xn(d), Cn(d * d), Mb(d * d), Q, wgt(d);
        for (int i = 0; i < d; i++) xn[i] = mid(y[i]);
        for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) {
            ival a(0.0), bq(0.0);
            for (int k = 0; k < d; k++) { if (C[k * d + j] != 0.0) a = a + J[i * d + k] * ival(C[k * d + j]); if (B[k * d + j] != 0.0) bq = bq + J[i * d + k] * ival(B[k * d + j]); }
            JC[i * d + j] = a; Cn[i * d + j] = mid(a);
            JB[i * d + j] = bq; Mb[i * d + j] = mid(bq);
        }
        for (int j = 0; j < d; j++) wgt[j] = width(r[j]);
        // QR-factorised B, block by block: the phi-block (rows 0..2n-1) and the W-block are orthogonalised
        // separately.  A single d x d QR mixes components whose scales differ by 1e6 (|phi| ~ 1e-6 next to
        // |W| ~ 1 near phi(-L_-)) and leaks the W rounding noise into phi, where it is amplified by e^{0.7 t}.
        std::vector<ival> Qi(d * d, ival(0.0));
        Q.assign(d * d, 0.0);
        bool qok = true;
        for (int blk = 0; blk < 2 && qok; blk++) {
            const int o = blk * dim;
            std::vector<double> Ms(dim * dim), ws(dim), Qs;
            std::vector<ival> Qis;
            for (int i = 0; i < dim; i++) { ws[i] = wgt[o + i]; for (int j = 0; j < dim; j++) Ms[i * dim + j] = Mb[(o + i) * d + o + j]; }
            qok = qr_orth(dim, Ms, ws, Qs) && inverse_orth(dim, Qs, Qis);
            if (qok) for (int i = 0; i < dim; i++) for (int j = 0; j < dim; j++) { Q[(o + i) * d + o + j] = Qs[i * dim + j]; Qi[(o + i) * d + o + j] = Qis[i * dim + j]; }
        }
        if (!qok) { Q.assign(d * d, 0.0); Qi.assign(d * d, ival(0.0)); for (int i = 0; i < d; i++) { Q[i * d + i] = 1.0; Qi[i * d + i] = ival(1.0); } }
        // NB: qr_orth permutes columns; r is re-expressed in the new basis through Qi*JB, so no bookkeeping is needed.
        for (int i = 0; i < d; i++) {
            ival z = y[i] - ival(xn[i]);
            for (int j = 0; j < d; j++) z = z + (JC[i * d + j] - ival(Cn[i * d + j])) * r0[j];
            Z[i] = z;
        }
        for (int i = 0; i < d; i++) {
            ival a(0.0);
            for (int j = 0; j < d; j++) {
                ival m(0.0);
                for (int k = 0; k < d; k++) m = m + Qi[i * d + k] * JB[k * d + j];
                a = a + m * r[j];
            }
            for (int k = 0; k < d; k++) a = a + Qi[i * d + k] * Z[k];
            rn[i] = a;
        }
        x = xn; C = Cn; B = Q; r = rn;
        prevTime = prevTime + stepMade;
    }
}

// propagateManifold::frameDet minus the host-side L_+/below-L_- checks  (propagateManifold.cpp:153-186)
static inline void run_front(const ccp_front_desc& D, ccp_front_result& res, std::vector<ccp_series_rec>* series, int threads) {
    std::memset(&res, 0, sizeof(res));
    res.first_failure = -1;
    const int n = D.n, dim = 2 * n, p = D.order, g = D.grid;
    if (n < 2 || n > MAXN || p < 2 || p > CCP_MAX_ORDER || g < 1 || g > CCP_MAX_GRID || D.step_log2 < 1 || D.step_log2 > 20 || !(D.L_minus > 0) || !std::isfinite(D.L_minus) || std::ldexp(D.L_minus, D.step_log2) > 16777216.0) {
        res.status = CCP_ST_BAD_DESC; return;
    }
    ival b[MAXN], c[MAXN];
    for (int i = 0; i < n; i++) b[i] = ival(D.b[i].lo, D.b[i].hi);
    for (int i = 0; i < n - 1; i++) c[i] = ival(D.c[i].lo, D.c[i].hi);
    Tape T; build_tape(T, n, b, c);
    std::vector<Trajectory> traj(n);
    if (threads > 1) {
        std::vector<std::thread> th;
        for (int k = 0; k < n; k++) th.emplace_back([&, k]() { RoundUpGuard gd; total_trajectory(D, T, k, traj[k]); });
        for (auto& t : th) t.join();
    } else {
        for (int k = 0; k < n; k++) total_trajectory(D, T, k, traj[k]);
    }
    for (int k = 0; k < n; k++) if (traj[k].status != CCP_OK) { res.status = traj[k].status; return; }
    // topFrame::initialize (topFrame.cpp:4-17)
    const size_t len = traj[0].sub.size();
    for (int k = 1; k < n; k++) if (traj[k].sub.size() != len) { res.status = CCP_ST_BAD_DESC; return; }
    res.series_length = (int)len; res.steps = (int)(len / g);
    Counter cnt;
    for (size_t i = 0; i < len; i++) {
        Frame Fm[4];
        for (int ty = 0; ty < 4; ty++) for (int k = 0; k < n; k++) for (int m = 0; m < n; m++) Fm[ty][m][k] = traj[k].sub[i].U[m][ty];   // constructFrameSeries
        ival dL = detN(Fm[0], n), dR = detN(Fm[1], n), dV = detN(Fm[2], n);                                                          // constructDetSeries
        ival dD = jacobi_derivative(Fm[2], Fm[3], n);
        const ival time = traj[0].sub[i].time;
        count_step(cnt, (int)i, time, dL, dR, dV, dD);
        if (series) { ccp_series_rec rec; rec.time = {time.lo, time.hi}; rec.det_L = {dL.lo, dL.hi}; rec.det_R = {dR.lo, dR.hi}; rec.det_V = {dV.lo, dV.hi}; rec.det_D = {dD.lo, dD.hi}; series->push_back(rec); }
        if (i == 0) res.det_first = {dV.lo, dV.hi};
        if (i == len - 1) res.det_last = {dV.lo, dV.hi};
        if (!(finite(dL) && finite(dR) && finite(dV) && finite(dD))) res.status = CCP_ST_NONFINITE;
    }
    res.zero_count = cnt.zero_count; res.failure_count = cnt.failure_count; res.first_failure = cnt.first_failure; res.n_conj = cnt.n_conj;
    for (int i = 0; i < cnt.n_conj; i++) { res.conj_time[i] = cnt.conj_time[i]; res.conj_index[i] = cnt.conj_index[i]; }
    // getLastFrame (topFrame.cpp:265-299) and getFirstFrame data (topFrame.cpp:301-331)
    const int ldl = CCP_MAX_N + 2, ldf = CCP_MAX_N + 1;
    for (int k = 0; k < n; k++) for (int j = 0; j < dim; j++) {
        const ival& a = traj[k].last_full[(size_t)(j + dim) * 5 + 2];   res.last_frame[j * ldl + k] = {a.lo, a.hi};
        const ival& f = traj[k].first_full[(size_t)(j + dim) * 5 + 1];  res.first_frame[j * ldf + k] = {f.lo, f.hi};
    }
    for (int j = 0; j < dim; j++) {
        const ival& a = traj[n - 1].last_full[(size_t)j * 5 + 4];  res.last_frame[j * ldl + n] = {a.lo, a.hi};
        const ival& bq = traj[n - 1].last_full[(size_t)j * 5 + 2]; res.last_frame[j * ldl + n + 1] = {bq.lo, bq.hi};
        const ival& f = traj[0].first_full[(size_t)j * 5 + 4];     res.first_frame[j * ldf + n] = {f.lo, f.hi};
    }
}

} // namespace ref
} // namespace orc
