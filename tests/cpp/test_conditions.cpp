// Host-only test of host/ccp_manifold_conditions.hpp (mirror of localManifold::checkConditions and its helpers).  No GPU: the
// enclosure of DF over U that boundDFU would deliver is formed here over the WHOLE box (one sub-box) with the same formula.
#include "../../computing-conjugate-points_b200/host/ccp_manifold_conditions.hpp"
#include <cstdio>
#include <random>

using namespace ccp_host;
typedef long double ld;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static ld lambda_max_ld(int n, std::vector<std::vector<ld>> S) {          // cyclic Jacobi in long double
    for (int sweep = 0; sweep < 40; sweep++)
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
            if (S[p][q] == 0) continue;
            const ld th = (S[q][q] - S[p][p]) / (2 * S[p][q]);
            const ld t = (th >= 0 ? 1 : -1) / (fabsl(th) + sqrtl(th * th + 1)), c = 1 / sqrtl(t * t + 1), s = t * c;
            for (int k = 0; k < n; k++) { const ld a = S[k][p], b = S[k][q]; S[k][p] = c * a - s * b; S[k][q] = s * a + c * b; }
            for (int k = 0; k < n; k++) { const ld a = S[p][k], b = S[q][k]; S[p][k] = c * a - s * b; S[q][k] = s * a + c * b; }
        }
    ld m = S[0][0]; for (int i = 1; i < n; i++) m = std::max(m, S[i][i]);
    return m;
}

// Df(x) of the front field (derivative of the formula string, ODE_functions.cpp:222), interval evaluation
static IMatrix frontDf(int n, const IVector& par, const IVector& x) {
    IMatrix J(2 * n, IVector(2 * n, interval(0.0)));
    auto one_m2 = [&](int i) { return ia::sub(interval(1.0), ia::mul(interval(2.0), x[i])); };
    auto g = [&](int i) { return ia::mul(x[i], ia::sub(interval(1.0), x[i])); };
    auto w = [&](int i) { return ia::sub(x[i], interval(0.5)); };
    for (int i = 0; i < n; i++) {
        J[i][n + i] = interval(1.0);
        // d/du [u(u-.5)(1-u)] = 3u(1-u) - 1/2
        interval d = ia::neg(ia::mul(par[i], ia::sub(ia::mul(interval(3.0), g(i)), interval(0.5))));
        if (i > 0) { d = ia::sub(d, ia::mul(par[n + i - 1], g(i - 1))); J[n + i][i - 1] = ia::neg(ia::mul(par[n + i - 1], ia::mul(one_m2(i - 1), w(i)))); }
        if (i + 1 < n) { d = ia::sub(d, ia::mul(par[n + i], g(i + 1))); J[n + i][i + 1] = ia::neg(ia::mul(par[n + i], ia::mul(one_m2(i + 1), w(i)))); }
        J[n + i][i] = d;
    }
    return J;
}

int main() {
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> uni(-1, 1);
    // ---- maxEigenvalueSym / euclNorm / euclLNorm / ml against long double
    for (int trial = 0; trial < 200; trial++) {
        const int n = 2 + trial % 3;
        const double shift = (trial % 4 == 0) ? -5.0 : (trial % 4 == 1 ? 5.0 : 0.0), rad = (trial % 5 == 0) ? 1e-6 : 0.0;
        IMatrix A(n, IVector(n)); std::vector<std::vector<ld>> Sm(n, std::vector<ld>(n)), Am(n, std::vector<ld>(n));
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { const double v = uni(rng) + (i == j ? shift : 0.0); A[i][j] = interval(v - rad, v + rad); Am[i][j] = v; }
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Sm[i][j] = (Am[i][j] + Am[j][i]) / 2;
        const ld lmax = lambda_max_ld(n, Sm);
        const double l = euclLNorm(A).hi;
        CHECK((ld)l >= lmax && (ld)l - lmax <= 1e-12L + 4 * n * rad);
        std::vector<std::vector<ld>> Sn = Sm; for (auto& r : Sn) for (auto& x : r) x = -x;
        const ld lmin = -lambda_max_ld(n, Sn);
        const double m = ml(A).lo;
        CHECK((ld)m <= lmin && lmin - (ld)m <= 1e-12L + 4 * n * rad);
        std::vector<std::vector<ld>> G(n, std::vector<ld>(n, 0));
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) G[i][j] += Am[k][i] * Am[k][j];
        const ld two = sqrtl(lambda_max_ld(n, G));
        const double e = euclNorm(A).hi;
        CHECK((ld)e >= two && (ld)e - two <= 1e-11L + 8 * n * rad);
    }
    {
        const IVector v = {interval(3.0), interval(-4.0, -4.0), interval(0.0)};
        const interval nv = euclNorm(v);
        CHECK(nv.lo == nv.hi && nv.hi >= 5.0 && nv.hi < 5.0 + 1e-14);
    }
    {   // non-finite entries must never certify a bound: lambda_max = +inf  =>  euclLNorm +inf, ml -inf (rate / L_+ conditions then fail)
        const double bad[3] = {std::nan(""), INFINITY, -INFINITY};
        for (double b : bad) {
            IMatrix A = {{interval(1.0), interval(0.25)}, {interval(0.25), interval(b)}};
            CHECK(maxEigenvalueSym(A) == INFINITY && euclLNorm(A).hi == INFINITY && ml(A).lo == -INFINITY);
            IMatrix B = {{interval(1.0), interval(b)}, {interval(0.5), interval(2.0)}};
            CHECK(maxEigenvalueSym(B) == INFINITY && !(euclNorm(B).hi < INFINITY));
        }
    }
    std::printf("norm bounds ok\n");

    // ---- the n = 3 defaults at the equilibrium 0: Df(0) = [[0, I], [M, 0]], M = 1/2 [[b1,c12,0],[c12,b2,c23],[0,c23,b3]]
    const int n = 3, d = 6;
    const IVector par = {interval(1.0), interval(0.975), interval(0.95), interval(0.05), interval(-0.05)};
    IVector zero(d, interval(0.0));
    {
        const IVector f0 = frontField(n, par, zero);
        for (int i = 0; i < d; i++) CHECK(f0[i].lo <= 0 && f0[i].hi >= 0 && f0[i].hi - f0[i].lo < 1e-300);   // outward rounding leaves denormals
        IVector x(d); std::vector<double> xd(d);
        for (int i = 0; i < d; i++) { xd[i] = 0.5 + 0.4 * uni(rng); x[i] = interval(xd[i]); }
        const IVector f = frontField(n, par, x);
        const double b[3] = {1.0, 0.975, 0.95}, c12 = 0.05, c23 = -0.05;
        auto ff = [&](double u) { return u * (u - .5) * (1 - u); };
        auto gg = [&](double ui, double uj) { return uj * (1 - uj) * (ui - .5); };
        const double e[3] = {-b[0] * ff(xd[0]) - c12 * gg(xd[0], xd[1]), -b[1] * ff(xd[1]) - c12 * gg(xd[1], xd[0]) - c23 * gg(xd[1], xd[2]), -b[2] * ff(xd[2]) - c23 * gg(xd[2], xd[1])};
        for (int i = 0; i < n; i++) { CHECK(f[i].lo == xd[n + i] && f[i].hi == xd[n + i]); CHECK(f[n + i].lo <= e[i] && e[i] <= f[n + i].hi && f[n + i].hi - f[n + i].lo < 1e-14); }
        // finite-difference check of the derivative used below
        const IMatrix J = frontDf(n, par, x);
        for (int j = 0; j < n; j++) {
            IVector xp = x, xm = x; xp[j] = interval(xd[j] + 1e-6); xm[j] = interval(xd[j] - 1e-6);
            const IVector fp = frontField(n, par, xp), fm = frontField(n, par, xm);
            for (int i = 0; i < n; i++) CHECK(std::fabs((fp[n + i].lo - fm[n + i].lo) / 2e-6 - ia::mid(J[n + i][j])) < 1e-8);
        }
    }
    // eigen-coordinates: columns (x_i, +-sqrt(mu_i) x_i); unstable block first
    DMatrix Mh = {{0.5, 0.025, 0.0}, {0.025, 0.4875, -0.025}, {0.0, -0.025, 0.475}}, X;
    jacobiEigenvectors(Mh, X);
    DMatrix A(d, std::vector<double>(d));
    for (int j = 0; j < n; j++) {
        double mu = 0; for (int a = 0; a < n; a++) for (int b = 0; b < n; b++) mu += X[a][j] * Mh[a][b] * X[b][j];
        const double l = std::sqrt(mu);
        for (int k = 0; k < n; k++) { A[k][j] = X[k][j]; A[n + k][j] = l * X[k][j]; A[k][n + j] = X[k][j]; A[n + k][n + j] = -l * X[k][j]; }
    }
    localManifold Mf;
    Mf.dimension = d; Mf.stable = false; Mf.L = interval(0.05); Mf.params = par;
    Mf.F.A = A; Mf.F.p = zero;
    {
        const DMatrix Ai = approxInverse(A);             // a point matrix stands in for gaussInverseMatrix(A) in this test
        Mf.F.Ainv.assign(d, IVector(d));
        for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) Mf.F.Ainv[i][j] = interval(Ai[i][j] - 1e-15, Ai[i][j] + 1e-15);
    }
    auto DF_over = [&](const localManifold& m, const IVector& U) {       // Ainv * Df(p + A U) * A over the whole box
        IVector x(d);
        for (int i = 0; i < d; i++) { interval a = m.F.p[i]; for (int j = 0; j < d; j++) a = ia::add(a, ia::mul(interval(m.F.A[i][j]), U[j])); x[i] = a; }
        const IMatrix J = frontDf(n, par, x);
        IMatrix T(d, IVector(d)), R(d, IVector(d));
        for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) { interval a(0.0); for (int k = 0; k < d; k++) a = ia::add(a, ia::mul(J[i][k], interval(m.F.A[k][j]))); T[i][j] = a; }
        for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) { interval a(0.0); for (int k = 0; k < d; k++) a = ia::add(a, ia::mul(m.F.Ainv[i][k], T[k][j])); R[i][j] = a; }
        return R;
    };
    {
        Mf.U_flat_global = IVector(n, interval(-1e-3, 1e-3));
        const IVector U = Mf.constructU(Mf.U_flat_global);
        for (int i = 0; i < n; i++) { CHECK(U[i].lo == -1e-3 && U[i].hi == 1e-3); CHECK(U[n + i].hi > 0.05 * 1.7e-3 && U[n + i].hi < 0.05 * 1.74e-3 && U[n + i].lo == -U[n + i].hi); }
        const IMatrix DFU = DF_over(Mf, U);
        CHECK(Mf.checkIsolatingBlock(DFU, U));
        CHECK(Mf.checkRateCondition(DFU));
        CHECK(Mf.xi.lo > 0.6 && Mf.xi.hi < 0.68 && Mf.mu.hi < -0.3 && Mf.mu.lo > -0.68);     // slowest rate ~0.67; mu = l(Fyy) + |Fyx|/L with |Fyx| = O(r)
        CHECK(Mf.checkConditionsOn(DFU, U));
        localManifold Ms = Mf; Ms.stable = true;
        const IVector Us = Ms.constructU(Ms.U_flat_global);
        for (int i = 0; i < n; i++) CHECK(Us[n + i].lo == -1e-3 && Us[i].hi == U[n + i].hi);
        CHECK(Ms.checkConditionsOn(DF_over(Ms, Us), Us));
        std::printf("checkConditions ok: xi in [%.4f, %.4f], mu <= %.4f\n", Mf.xi.lo, Mf.xi.hi, Mf.mu.hi);
    }
    {
        localManifold Mb = Mf; Mb.U_flat_global = IVector(n, interval(-0.6, 0.6));        // far too large a neighbourhood
        const IVector U = Mb.constructU(Mb.U_flat_global);
        CHECK(!Mb.checkConditionsOn(DF_over(Mb, U), U));
        localManifold Mc = Mf; Mc.L = interval(1e-7);                                      // cone too thin for the coupling: rate condition fails
        Mc.U_flat_global = IVector(n, interval(-1e-2, 1e-2));
        const IVector Uc = Mc.constructU(Mc.U_flat_global);
        CHECK(!Mc.checkRateCondition(DF_over(Mc, Uc)));
        std::printf("failing neighbourhoods rejected\n");
    }
    {
        const IVector r = Mf.containmentRatios({interval(5e-4), interval(-2.5e-4), interval(1e-3)});
        CHECK(std::fabs(ia::mid(r[0]) - 0.5) < 1e-12 && std::fabs(ia::mid(r[1]) - 0.25) < 1e-12 && r[2].lo <= 1.0 && r[2].hi >= 1.0);
        const IMatrix Z = {{interval(1.0), interval(2.0)}, {interval(3.0), interval(4.0)}};
        const std::vector<IMatrix> b = blockDecompose(Z, 2);
        CHECK(b[0][0][0].lo == 1 && b[1][0][0].lo == 2 && b[2][0][0].lo == 3 && b[3][0][0].lo == 4);
        CHECK(getMax({interval(-1.0, 0.5), interval(-3.0, 0.75), interval(0.1, 0.2)}).lo == 0.75);
        std::printf("helpers ok\n");
    }
    return 0;
}
