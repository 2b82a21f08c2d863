// ccp_setup.cpp -- host-side generation of front descriptors (plain doubles, NOT rigorous).
//
// The hot path (frameDet) consumes the output of stage (i) of the reference pipeline:
// the point phi(-L_-) on the local unstable manifold, the eigenbasis A_u and the error boxes.
// The reference computes them with CAPD (heteroclinic.cpp:96-330); CAPD is not available, so
// this file restates the *numerical* part of that stage in doubles, following the reference's
// own recipe step by step, to feed synthetic sweeps:
//   coordinateChange + symplecticNormalization   ODE_functions.cpp:18-50, utils.cpp:110-124,456-495
//   approxT                                      ODE_functions.cpp:360-375
//   initialGuessGlobal / knownSolution           ODE_functions.cpp:52-64,101-121
//   projectPoint / containmentRatios rescale     localManifold.cpp:105-132,333-344; ODE_functions.cpp:278-305
//   FindTime                                     boundaryValueProblem.cpp:118-215
//   NewtonStep (zero-width, 20x)                 boundaryValueProblem.cpp:237-317, heteroclinic.cpp:186-191
//   L_minus = sup(2*T*0.84)                      heteroclinic.cpp:72,324
//   getPointNbd / constructU                     localManifold.cpp:4-41,73-102
// Error-box magnitudes that the reference obtains from its validation stages (Krawczyk
// neighbourhood, eigenfunction error) are inputs here (ccp_setup_opts), see SURVEY.md 8(d).
#include "../../include/ccp_setup.h"
#include <cmath>
#include <cstring>
#include <cfloat>
#include <thread>
#include <vector>
#include <atomic>
#include <algorithm>

namespace {

constexpr int MAXN = CCP_MAX_N;
constexpr int MAXD = CCP_MAX_DIM;
constexpr int ORD  = 20;          // float integrator order
constexpr double HSTEP = 0.5;     // float integrator step (solution is analytic in |Im t| < pi/lambda ~ 4.4)

struct Sys { int n; double b[MAXN]; double c[MAXN]; };   // c[i] couples i and i+1

// f(phi): phi = (u_1..u_n, v_1..v_n)
static void field(const Sys& s, const double* phi, double* out) {
    int n = s.n;
    double g[MAXN], w[MAXN];
    for (int i = 0; i < n; i++) { w[i] = phi[i] - 0.5; g[i] = phi[i] * (1.0 - phi[i]); }
    for (int i = 0; i < n; i++) {
        out[i] = phi[n + i];
        double a = -s.b[i] * g[i] * w[i];
        if (i > 0)     a -= s.c[i - 1] * g[i - 1] * w[i];
        if (i < n - 1) a -= s.c[i] * g[i + 1] * w[i];
        out[n + i] = a;
    }
}
// M(u) = D_u (v')
static void hess(const Sys& s, const double* u, double M[MAXN][MAXN]) {
    int n = s.n;
    double g[MAXN], w[MAXN];
    for (int i = 0; i < n; i++) { w[i] = u[i] - 0.5; g[i] = u[i] * (1.0 - u[i]); }
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) M[i][j] = 0.0;
    for (int i = 0; i < n; i++) {
        double d = -s.b[i] * (3.0 * g[i] - 0.5);
        if (i > 0)     d -= s.c[i - 1] * g[i - 1];
        if (i < n - 1) d -= s.c[i] * g[i + 1];
        M[i][i] = d;
        if (i < n - 1) { double o = 2.0 * s.c[i] * w[i] * w[i + 1]; M[i][i + 1] = o; M[i + 1][i] = o; }
    }
}

// One Taylor step of phi (2n) and optionally Psi (2n x 2n, row-major ld 2n) of size h (may be < 0).
// Built for AVX2 too where the host has it (function multi-versioning, resolved at load time).  The AVX2 target has no fused
// multiply-add, and nothing here may be contracted or re-associated: the descriptors must not depend on the host.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx2", "default")))
#endif
static void taylor_step(const Sys& s, double* phi, double* Psi, double h) {
    const int n = s.n, D = 2 * n;
    static thread_local double u[ORD + 2][MAXN], v[ORD + 2][MAXN], w[ORD + 2][MAXN], g[ORD + 2][MAXN], q[ORD + 2][MAXN];
    static thread_local double Md[ORD + 2][MAXN], Mo[ORD + 2][MAXN];
    static thread_local double Pu[ORD + 2][MAXN][MAXD], Pv[ORD + 2][MAXN][MAXD];
    for (int i = 0; i < n; i++) { u[0][i] = phi[i]; v[0][i] = phi[n + i]; }
    if (Psi) for (int i = 0; i < n; i++) for (int cidx = 0; cidx < D; cidx++) {
        Pu[0][i][cidx] = Psi[i * D + cidx]; Pv[0][i][cidx] = Psi[(n + i) * D + cidx];
    }
    for (int k = 0; k <= ORD; k++) {
        for (int i = 0; i < n; i++) w[k][i] = u[k][i] - (k == 0 ? 0.5 : 0.0);
        // The Cauchy sums of one order side by side (full MAXN-wide rows; components >= n are never written and only feed sums nobody
        // reads): every sum sees its terms in the same order as a sum-by-sum loop -- the same bits --, but the chains are independent
        // of each other instead of one latency-bound chain after the other
        {
            double ag[MAXN], aq[MAXN];
            for (int i = 0; i < MAXN; i++) { ag[i] = 0; aq[i] = 0; }
            for (int j = 0; j <= k; j++) {
                const double* wj = w[j]; const double* wr = w[k - j];
                for (int i = 0; i < MAXN; i++) ag[i] += wj[i] * wr[i];
                for (int i = 0; i + 1 < MAXN; i++) aq[i] += wj[i] * wr[i + 1];
            }
            for (int i = 0; i < n; i++) g[k][i] = (k == 0 ? 0.25 : 0.0) - ag[i];
            for (int i = 0; i + 1 < n; i++) q[k][i] = aq[i];
            double sg[MAXN], xm[MAXN], xp[MAXN];
            for (int i = 0; i < MAXN; i++) { sg[i] = 0; xm[i] = 0; xp[i] = 0; }
            for (int j = 0; j <= k; j++) {
                const double* gj = g[j]; const double* wr = w[k - j];
                for (int i = 0; i < MAXN; i++) sg[i] += gj[i] * wr[i];
                for (int i = 1; i < MAXN; i++) xm[i] += gj[i - 1] * wr[i];
                for (int i = 0; i + 1 < MAXN; i++) xp[i] += gj[i + 1] * wr[i];
            }
            for (int i = 0; i < n; i++) {
                double f = -s.b[i] * sg[i];
                if (i > 0)     f -= s.c[i - 1] * xm[i];
                if (i < n - 1) f -= s.c[i] * xp[i];
                u[k + 1][i] = v[k][i] / (k + 1);
                v[k + 1][i] = f / (k + 1);
            }
        }
        if (Psi) {
            for (int i = 0; i < n; i++) {
                double d = -s.b[i] * (3.0 * g[k][i] - (k == 0 ? 0.5 : 0.0));
                if (i > 0)     d -= s.c[i - 1] * g[k][i - 1];
                if (i < n - 1) d -= s.c[i] * g[k][i + 1];
                Md[k][i] = d;
                if (i < n - 1) Mo[k][i] = 2.0 * s.c[i] * q[k][i];
            }
            // all MAXD columns of a row at once (the columns >= D stay zero): every element sees the same operations in the same order as
            // a column-by-column loop -- the same bits --, but the inner loops have a fixed length and vectorise
            for (int i = 0; i < n; i++) {
                double a[MAXD];
                for (int cidx = 0; cidx < MAXD; cidx++) a[cidx] = 0;
                for (int j = 0; j <= k; j++) {
                    const double md = Md[j][i];
                    const double* p0 = Pu[k - j][i];
                    for (int cidx = 0; cidx < MAXD; cidx++) a[cidx] += md * p0[cidx];
                    if (i > 0)     { const double mo = Mo[j][i - 1]; const double* p1 = Pu[k - j][i - 1]; for (int cidx = 0; cidx < MAXD; cidx++) a[cidx] += mo * p1[cidx]; }
                    if (i < n - 1) { const double mo = Mo[j][i];     const double* p2 = Pu[k - j][i + 1]; for (int cidx = 0; cidx < MAXD; cidx++) a[cidx] += mo * p2[cidx]; }
                }
                const double kk = k + 1;
                for (int cidx = 0; cidx < MAXD; cidx++) { Pu[k + 1][i][cidx] = Pv[k][i][cidx] / kk; Pv[k + 1][i][cidx] = a[cidx] / kk; }
            }
        }
    }
    for (int i = 0; i < n; i++) {
        double su = 0, sv = 0;
        for (int k = ORD; k >= 0; k--) { su = su * h + u[k][i]; sv = sv * h + v[k][i]; }
        phi[i] = su; phi[n + i] = sv;
    }
    if (Psi) for (int i = 0; i < n; i++) {
        double su[MAXD], sv[MAXD];
        for (int cidx = 0; cidx < MAXD; cidx++) { su[cidx] = 0; sv[cidx] = 0; }
        for (int k = ORD; k >= 0; k--) for (int cidx = 0; cidx < MAXD; cidx++) { su[cidx] = su[cidx] * h + Pu[k][i][cidx]; sv[cidx] = sv[cidx] * h + Pv[k][i][cidx]; }
        for (int cidx = 0; cidx < D; cidx++) { Psi[i * D + cidx] = su[cidx]; Psi[(n + i) * D + cidx] = sv[cidx]; }
    }
}

// flow for time T (T may be negative), Psi optional (initialised to identity here)
static void flow(const Sys& s, double* phi, double* Psi, double T) {
    const int D = 2 * s.n;
    if (Psi) { for (int i = 0; i < D * D; i++) Psi[i] = 0; for (int i = 0; i < D; i++) Psi[i * D + i] = 1; }
    double sgn = T < 0 ? -1.0 : 1.0, left = std::fabs(T);
    while (left > 0) {
        double h = std::min(HSTEP, left);
        taylor_step(s, phi, Psi, sgn * h);
        left -= h;
    }
}

// symmetric Jacobi eigen-decomposition (n<=4): K = V diag(mu) V^T
static void jacobi_eig(int n, double K[MAXN][MAXN], double mu[MAXN], double V[MAXN][MAXN]) {
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i][j] = (i == j);
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0; for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += K[i][j] * K[i][j];
        if (off < 1e-300) break;
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
            if (K[p][q] == 0.0) continue;
            double th = (K[q][q] - K[p][p]) / (2.0 * K[p][q]);
            double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
            double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
            for (int k = 0; k < n; k++) { double a = K[k][p], b = K[k][q]; K[k][p] = cs * a - sn * b; K[k][q] = sn * a + cs * b; }
            for (int k = 0; k < n; k++) { double a = K[p][k], b = K[q][k]; K[p][k] = cs * a - sn * b; K[q][k] = sn * a + cs * b; }
            for (int k = 0; k < n; k++) { double a = V[k][p], b = V[k][q]; V[k][p] = cs * a - sn * b; V[k][q] = sn * a + cs * b; }
        }
    }
    for (int i = 0; i < n; i++) mu[i] = K[i][i];
}

// Eigenbasis of Df at an equilibrium, sorted by eigenvalue descending, symplectically normalised.
// At both (0,..,0) and (1,..,1): Df = [[0,I],[K,0]], K = tridiag(b_i/2 ; c_ij/2)  (g=0, w=+-1/2).
static void eigenbasis(const Sys& s, double A[MAXD][MAXD], double lam[MAXN]) {
    const int n = s.n, D = 2 * n;
    double K[MAXN][MAXN] = {}, mu[MAXN], V[MAXN][MAXN];
    for (int i = 0; i < n; i++) { K[i][i] = 0.5 * s.b[i]; if (i < n - 1) { K[i][i + 1] = 0.5 * s.c[i]; K[i + 1][i] = 0.5 * s.c[i]; } }
    jacobi_eig(n, K, mu, V);
    int idx[MAXN]; for (int i = 0; i < n; i++) idx[i] = i;
    std::sort(idx, idx + n, [&](int a, int b) { return mu[a] > mu[b]; });
    for (int j = 0; j < n; j++) {
        int e = idx[j];
        double l = std::sqrt(mu[e]); lam[j] = l;
        // sign convention: largest-magnitude component of the u-part positive
        int big = 0; for (int i = 1; i < n; i++) if (std::fabs(V[i][e]) > std::fabs(V[big][e])) big = i;
        double sg = V[big][e] < 0 ? -1.0 : 1.0;
        double nrm = 0; for (int i = 0; i < n; i++) nrm += V[i][e] * V[i][e]; nrm = std::sqrt(nrm);
        double sc = sg / nrm / std::sqrt(2.0 * l);          // omega(p,q) = 2*lambda after unit-normalising the top part
        for (int i = 0; i < n; i++) {
            A[i][j] = V[i][e] * sc;          A[n + i][j] = l * V[i][e] * sc;
            A[i][D - 1 - j] = V[i][e] * sc;  A[n + i][D - 1 - j] = -l * V[i][e] * sc;
        }
    }
    // The eigenvector signs returned by the reference's alglib call are not recoverable; fix the one free
    // global sign so that det(top n x n unstable block) > 0, which is what plot_det.txt shows at t = 0.
    double Mx[MAXN * MAXN]; double det = 1.0;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Mx[i * n + j] = A[i][j];
    for (int k = 0; k < n; k++) {
        int piv = k; for (int i = k + 1; i < n; i++) if (std::fabs(Mx[i * n + k]) > std::fabs(Mx[piv * n + k])) piv = i;
        if (piv != k) { for (int j = 0; j < n; j++) std::swap(Mx[k * n + j], Mx[piv * n + j]); det = -det; }
        det *= Mx[k * n + k];
        for (int i = k + 1; i < n; i++) { double f = Mx[i * n + k] / Mx[k * n + k]; for (int j = k; j < n; j++) Mx[i * n + j] -= f * Mx[k * n + j]; }
    }
    if (det < 0) for (int i = 0; i < D; i++) { A[i][0] = -A[i][0]; A[i][D - 1] = -A[i][D - 1]; }
}

static bool solve_linear(int D, double* Mx, double* rhs) {   // Gaussian elimination, partial pivoting, in place
    for (int k = 0; k < D; k++) {
        int piv = k; for (int i = k + 1; i < D; i++) if (std::fabs(Mx[i * D + k]) > std::fabs(Mx[piv * D + k])) piv = i;
        if (Mx[piv * D + k] == 0.0) return false;
        if (piv != k) { for (int j = 0; j < D; j++) std::swap(Mx[k * D + j], Mx[piv * D + j]); std::swap(rhs[k], rhs[piv]); }
        for (int i = k + 1; i < D; i++) {
            double f = Mx[i * D + k] / Mx[k * D + k];
            for (int j = k; j < D; j++) Mx[i * D + j] -= f * Mx[k * D + j];
            rhs[i] -= f * rhs[k];
        }
    }
    for (int k = D - 1; k >= 0; k--) {
        double a = rhs[k]; for (int j = k + 1; j < D; j++) a -= Mx[k * D + j] * rhs[j];
        rhs[k] = a / Mx[k * D + k];
    }
    return true;
}

static inline double up(double x) { return std::nextafter(x, INFINITY); }
static inline double dn(double x) { return std::nextafter(x, -INFINITY); }

static int setup_one(int n, const double* params, const ccp_setup_opts* o, ccp_front_desc* out, ccp_setup_info* info) {
    if (n < 2 || n > MAXN) return CCP_E_ARG;
    const int D = 2 * n;
    Sys s; s.n = n;
    for (int i = 0; i < n; i++) s.b[i] = params[i];
    for (int i = 0; i < n - 1; i++) s.c[i] = params[n + i];
    double scale  = (o && o->scale > 0) ? o->scale : (n == 2 ? 0.000016 : 0.000001);        // heteroclinic.cpp:80-87
    double coneL  = (o && o->cone_L > 0) ? o->cone_L : (n == 2 ? 0.00008 : 0.000015);
    double eigerr = (o && o->eig_err >= 0) ? o->eig_err : (n == 2 ? 2e-4 : 2e-5);
    double xynbd  = (o && o->xy_nbd >= 0) ? o->xy_nbd : 1e-12;              // XY_nbd x-part +-1e-12
    double pct    = (o && o->L_minus_percent > 0) ? o->L_minus_percent : 0.84;
    int newton    = (o && o->newton_steps > 0) ? o->newton_steps : 20;

    double Au[MAXD][MAXD] = {}, As[MAXD][MAXD] = {}, lam[MAXN];
    eigenbasis(s, Au, lam);
    eigenbasis(s, As, lam);          // K is identical at both equilibria for this family

    // approxT (ODE_functions.cpp:360-375)
    double bmin = s.b[0]; for (int i = 1; i < n; i++) bmin = std::min(bmin, s.b[i]);
    double T = std::log(1.0 / scale - 1.0) / std::sqrt(bmin / 2.0);

    // initial guesses from the uncoupled front at -T (unstable side) and +T (stable side)
    double gU[MAXD], gS[MAXD];
    for (int i = 0; i < n; i++) {
        double rt = std::sqrt(s.b[i] / 2.0);
        gU[i] = 1.0 / (1.0 + std::exp(rt * T));  gU[n + i] = rt / (2.0 + std::exp(-rt * T) + std::exp(rt * T));
        gS[i] = 1.0 / (1.0 + std::exp(-rt * T)); gS[n + i] = rt / (2.0 + std::exp(rt * T) + std::exp(-rt * T));
        gS[i] -= 1.0;                              // projectPoint subtracts the equilibrium (1,..,1|0)
    }
    double x[MAXN], y[MAXN];
    for (int i = 0; i < n; i++) {                  // successive deflation, localManifold.cpp:116-130
        double a = 0; for (int r = 0; r < D; r++) a += Au[r][i] * gU[r];
        x[i] = a; for (int r = 0; r < D; r++) gU[r] -= a * Au[r][i];
    }
    for (int i = 0; i < n; i++) {
        double a = 0; for (int r = 0; r < D; r++) a += As[r][n + i] * gS[r];
        y[i] = a; for (int r = 0; r < D; r++) gS[r] -= a * As[r][n + i];
    }
    double mx = 0, my = 0;
    for (int i = 0; i < n; i++) { mx = std::max(mx, std::fabs(x[i]) / scale); my = std::max(my, std::fabs(y[i]) / scale); }
    for (int i = 0; i < n; i++) { x[i] /= mx; y[i] /= my; }
    double r_u_sqr = 0; for (int i = 0; i < n; i++) r_u_sqr += x[i] * x[i];

    auto endpoints = [&](double Tt, double* Gx, double* Gy, double* Px, double* Py) {
        for (int r = 0; r < D; r++) { double a = 0, bq = 0; for (int i = 0; i < n; i++) { a += Au[r][i] * x[i]; bq += As[r][n + i] * y[i]; } Gx[r] = a; Gy[r] = bq + (r < n ? 1.0 : 0.0); }
        flow(s, Gx, Px, Tt);
        flow(s, Gy, Py, -Tt);
    };

    // FindTime: 10 Newton steps on T minimising |Phi_T(x)-Phi_-T(y)|^2/2 (boundaryValueProblem.cpp:118-215)
    {
        double T0 = T, d0_first = 0, d0 = 0;
        for (int it = 0; it < 10; it++) {
            double Gx[MAXD], Gy[MAXD], fx[MAXD], fy[MAXD];
            endpoints(T, Gx, Gy, nullptr, nullptr);
            field(s, Gx, fx); field(s, Gy, fy);
            double Mx[MAXN][MAXN], My[MAXN][MAXN];
            hess(s, Gx, Mx); hess(s, Gy, My);
            double ddx[MAXD], ddy[MAXD];      // Df(G) f(G)
            for (int i = 0; i < n; i++) {
                ddx[i] = fx[n + i]; ddy[i] = fy[n + i];
                double a = 0, bq = 0; for (int j = 0; j < n; j++) { a += Mx[i][j] * fx[j]; bq += My[i][j] * fy[j]; }
                ddx[n + i] = a; ddy[n + i] = bq;
            }
            double D0 = 0, D1 = 0, D2 = 0;
            for (int r = 0; r < D; r++) {
                double dv = Gx[r] - Gy[r];
                double d1 = fx[r] + fy[r];              // phi_T_y_deriv = -f(Gy)
                double d2 = ddx[r] - ddy[r];            // phi_T_y_DD*phi_T_y_deriv = (-Df)(-f) = Df f
                D0 += dv * dv; D1 += dv * d1; D2 += d1 * d1 + dv * d2;
            }
            if (it == 0) d0_first = D0;
            d0 = D0;
            T = T - D1 / D2;
        }
        if (d0_first < d0) T = T0;
    }

    // zero-width Newton on (x,y), last equation |x|^2 = r_u^2 (boundaryValueProblem.cpp:237-353).
    // The reference takes 20 plain Newton steps; for weak couplings the phase-locking directions are soft and
    // plain Newton from the uncoupled guess can overshoot, so the step is damped by backtracking on |G|.
    double resid = 0; int converged = 0;
    auto residual = [&](const double* xx, const double* yy, double* G, double* Px, double* Py) {
        double Gx[MAXD], Gy[MAXD];
        for (int r = 0; r < D; r++) { double a = 0, bq = 0; for (int i = 0; i < n; i++) { a += Au[r][i] * xx[i]; bq += As[r][n + i] * yy[i]; } Gx[r] = a; Gy[r] = bq + (r < n ? 1.0 : 0.0); }
        flow(s, Gx, Px, T);
        flow(s, Gy, Py, -T);
        for (int r = 0; r < D; r++) G[r] = Gx[r] - Gy[r];
        double xr = 0; for (int i = 0; i < n; i++) xr += xx[i] * xx[i];
        G[D - 1] = (xr - r_u_sqr) / r_u_sqr;          // relative radius defect (same zero set as the reference's row)
        double nn = 0; for (int r = 0; r < D; r++) nn += G[r] * G[r];
        return std::sqrt(nn);
    };
    // Levenberg-Marquardt in scaled variables z = (x,y)/scale: the weak-coupling phase modes make DG nearly
    // singular away from the solution, where a pure Newton direction is dominated by the null vector.
    const int max_it = std::max(newton, 400);
    double G[MAXD], Px[MAXD * MAXD], Py[MAXD * MAXD];
    double gn = residual(x, y, G, Px, Py);
    double mu = 1e-2;
    for (int it = 0; it < max_it && std::isfinite(gn); it++) {
        double DG[MAXD * MAXD];
        for (int r = 0; r < D; r++) for (int j = 0; j < D; j++) {
            double a = 0;
            if (j < n) { for (int k = 0; k < D; k++) a += Px[r * D + k] * Au[k][j]; }
            else       { for (int k = 0; k < D; k++) a -= Py[r * D + k] * As[k][j]; }
            DG[r * D + j] = a * scale;
        }
        for (int j = 0; j < D; j++) DG[(D - 1) * D + j] = (j < n) ? 2.0 * x[j] / r_u_sqr * scale : 0.0;
        double JtJ[MAXD * MAXD], Jtg[MAXD];
        for (int i = 0; i < D; i++) {
            double a = 0; for (int r = 0; r < D; r++) a += DG[r * D + i] * G[r];
            Jtg[i] = a;
            for (int j = 0; j < D; j++) { double b2 = 0; for (int r = 0; r < D; r++) b2 += DG[r * D + i] * DG[r * D + j]; JtJ[i * D + j] = b2; }
        }
        bool accepted = false; double stepn = 0;
        for (int tries = 0; tries < 30 && !accepted; tries++) {
            double Mx[MAXD * MAXD], st[MAXD];
            for (int i = 0; i < D * D; i++) Mx[i] = JtJ[i];
            for (int i = 0; i < D; i++) { Mx[i * D + i] += mu * (JtJ[i * D + i] + 1e-30); st[i] = Jtg[i]; }
            if (!solve_linear(D, Mx, st)) { mu *= 10; continue; }
            double xn[MAXN], yn[MAXN], Gn[MAXD], Pxn[MAXD * MAXD], Pyn[MAXD * MAXD];
            for (int i = 0; i < n; i++) { xn[i] = x[i] - scale * st[i]; yn[i] = y[i] - scale * st[n + i]; }
            // the trial point is judged on the residual alone (flows without the variational columns: an eighth of the work, the same
            // values of G); only an accepted point pays for D Phi
            double gnew = residual(xn, yn, Gn, nullptr, nullptr);
            if (std::isfinite(gnew) && gnew < gn) {
                gnew = residual(xn, yn, Gn, Pxn, Pyn);
                stepn = 0; for (int i = 0; i < D; i++) stepn = std::max(stepn, std::fabs(st[i]));
                for (int i = 0; i < n; i++) { x[i] = xn[i]; y[i] = yn[i]; }
                for (int r = 0; r < D; r++) G[r] = Gn[r];
                for (int i = 0; i < D * D; i++) { Px[i] = Pxn[i]; Py[i] = Pyn[i]; }
                gn = gnew; mu = std::max(mu * 0.2, 1e-12); accepted = true;
            } else mu *= 5;
        }
        if (!accepted) break;                     // noise floor
        if (gn < 1e-9 && stepn < 1e-9) break;
    }
    resid = gn;
    converged = (std::isfinite(gn) && gn < 1e-8) ? 1 : 0;

    // ---- assemble the descriptor ----
    std::memset(out, 0, sizeof(*out));
    out->n = n;
    out->order = (o && o->order > 0) ? o->order : 20;
    out->step_log2 = (o && o->step_log2 > 0) ? o->step_log2 : 5;
    out->grid = (o && o->grid > 0) ? o->grid : 14;
    double Lm = (o && o->L_minus_override > 0) ? o->L_minus_override : up(up(2.0 * T) * pct);   // sup(2*T*0.84)
    out->L_minus = Lm;
    for (int i = 0; i < n; i++) { out->b[i].lo = out->b[i].hi = s.b[i]; }
    for (int i = 0; i < n - 1; i++) { out->c[i].lo = out->c[i].hi = s.c[i]; }
    for (int r = 0; r < D; r++) for (int j = 0; j < D; j++) out->A_u[r * CCP_MAX_DIM + j] = Au[r][j];
    double xnorm = 0;
    for (int i = 0; i < n; i++) { double a = std::fabs(x[i]) + xynbd; xnorm += a * a; }
    xnorm = up(std::sqrt(up(xnorm)));
    for (int r = 0; r < D; r++) {             // p_i = p_u + A_u*(x,0), outward-inflated dot product
        double a = 0, m = 0; for (int i = 0; i < n; i++) { a += Au[r][i] * x[i]; m += std::fabs(Au[r][i] * x[i]); }
        double e = m * (4.0 * n) * DBL_EPSILON;
        out->p_i[r].lo = dn(a - e); out->p_i[r].hi = up(a + e);
    }
    for (int i = 0; i < n; i++) {
        double r = up(xynbd * (1.0 + 4 * DBL_EPSILON));
        out->Uxy[i].lo = -r; out->Uxy[i].hi = r;
        double e = up(coneL * xnorm);          // +- L*||x+nbd||  (localManifold.cpp:12-20)
        out->Uxy[n + i].lo = -e; out->Uxy[n + i].hi = e;
    }
    for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) { out->Eu_err[j * CCP_MAX_N + k].lo = -eigerr; out->Eu_err[j * CCP_MAX_N + k].hi = eigerr; }
    if (info) {
        std::memset(info, 0, sizeof(*info));
        info->T = T; info->L_minus = Lm; info->newton_residual = resid; info->converged = converged;
        for (int i = 0; i < n; i++) { info->eig_unstable[i] = lam[i]; info->x_local[i] = x[i]; info->y_local[i] = y[i]; }
    }
    return 0;
}

} // namespace

extern "C" int ccp_setup_front(int n, const double* params, const ccp_setup_opts* opts, ccp_front_desc* out, ccp_setup_info* info) {
    if (!params || !out) return CCP_E_ARG;
    return setup_one(n, params, opts, out, info);
}

extern "C" int ccp_setup_sweep(int n, const double* params, size_t n_fronts, const ccp_setup_opts* opts,
                               ccp_front_desc* out, ccp_setup_info* info, int threads) {
    if (!params || !out || n < 2 || n > MAXN) return CCP_E_ARG;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if ((size_t)nt > n_fronts) nt = (int)std::max<size_t>(1, n_fronts);
    std::atomic<size_t> next(0);
    std::atomic<int> rc(0);
    const int np = 2 * n - 1;
    auto work = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n_fronts) break;
            int r = setup_one(n, params + i * np, opts, out + i, info ? info + i : nullptr);
            if (r) rc = r;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return rc;
}

extern "C" int ccp_parameter_list(int N_c, int N_r, double r_min, double r_max, const double* b3, double* out) {
    // Construct_ParameterList, heteroclinic.cpp:386-427
    if (N_c < 1 || N_r < 1 || !out) return CCP_E_ARG;
    double b[3] = {1.0, 0.975, 0.95};
    if (b3) { b[0] = b3[0]; b[1] = b3[1]; b[2] = b3[2]; }
    double radius_step = N_r > 1 ? (r_max - r_min) / (N_r - 1) : 0.0;
    size_t k = 0;
    for (int i = 0; i < N_c; i++) {
        double theta = (i + .5) * 2 * M_PI / N_c;
        for (int j = 0; j < N_r; j++) {
            double radius = r_min + radius_step * j;
            out[k * 5 + 0] = b[0]; out[k * 5 + 1] = b[1]; out[k * 5 + 2] = b[2];
            out[k * 5 + 3] = radius * std::cos(theta);
            out[k * 5 + 4] = radius * std::sin(theta);
            k++;
            theta = theta + M_PI / ((double)N_c * N_r);
        }
    }
    return 0;
}
