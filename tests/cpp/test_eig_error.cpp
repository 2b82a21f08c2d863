// Host-only test of host/ccp_local_manifold_eig.hpp (mirror of localManifold_Eig: the producer of the descriptor field Eu_err).
// Setting: the n = 3 defaults at the equilibrium 0, eigen-coordinates with the symplectic normalisation |pi_1 v|^2 = 1/(2|lambda|);
// DF over the neighbourhood is enclosed on the host over the whole box (what boundDFU delivers from the GPU, without subdivision).
#include "../../computing-conjugate-points_b200/host/ccp_local_manifold_eig.hpp"
#include <cstdio>

using namespace ccp_host;
typedef long double ld;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static void jacobi_ld(int n, std::vector<std::vector<ld>>& S) {
    for (int sweep = 0; sweep < 40; sweep++)
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
            if (S[p][q] == 0) continue;
            const ld th = (S[q][q] - S[p][p]) / (2 * S[p][q]);
            const ld t = (th >= 0 ? 1 : -1) / (fabsl(th) + sqrtl(th * th + 1)), c = 1 / sqrtl(t * t + 1), s = t * c;
            for (int k = 0; k < n; k++) { const ld a = S[k][p], b = S[k][q]; S[k][p] = c * a - s * b; S[k][q] = s * a + c * b; }
            for (int k = 0; k < n; k++) { const ld a = S[p][k], b = S[q][k]; S[p][k] = c * a - s * b; S[q][k] = s * a + c * b; }
        }
}
static ld two_norm_ld(const std::vector<std::vector<ld>>& A, bool smallest = false) {      // sqrt of the extreme eigenvalue of A^T A
    const int r = (int)A.size(), c = (int)A[0].size();
    std::vector<std::vector<ld>> G(c, std::vector<ld>(c, 0));
    for (int i = 0; i < c; i++) for (int j = 0; j < c; j++) for (int k = 0; k < r; k++) G[i][j] += A[k][i] * A[k][j];
    jacobi_ld(c, G);
    ld m = G[0][0]; for (int i = 1; i < c; i++) m = smallest ? std::min(m, G[i][i]) : std::max(m, G[i][i]);
    return sqrtl(m);
}

int main() {
    const int n = 3, d = 6;
    const IVector par = {interval(1.0), interval(0.975), interval(0.95), interval(0.05), interval(-0.05)};
    const IVector zero(d, interval(0.0));
    DMatrix Mh = {{0.5, 0.025, 0.0}, {0.025, 0.4875, -0.025}, {0.0, -0.025, 0.475}};
    // eigen-decomposition in long double (the columns must meet the 3e-14 start box of boundSingleEigenvector after rounding)
    std::vector<std::vector<ld>> S(n, std::vector<ld>(n)), X(n, std::vector<ld>(n, 0));
    for (int i = 0; i < n; i++) { X[i][i] = 1; for (int j = 0; j < n; j++) S[i][j] = Mh[i][j]; }
    for (int sweep = 0; sweep < 40; sweep++)
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
            if (S[p][q] == 0) continue;
            const ld th = (S[q][q] - S[p][p]) / (2 * S[p][q]);
            const ld t = (th >= 0 ? 1 : -1) / (fabsl(th) + sqrtl(th * th + 1)), c = 1 / sqrtl(t * t + 1), sn = t * c;
            for (int k = 0; k < n; k++) { const ld a = S[k][p], b = S[k][q]; S[k][p] = c * a - sn * b; S[k][q] = sn * a + c * b; }
            for (int k = 0; k < n; k++) { const ld a = S[p][k], b = S[q][k]; S[p][k] = c * a - sn * b; S[q][k] = sn * a + c * b; }
            for (int k = 0; k < n; k++) { const ld a = X[k][p], b = X[k][q]; X[k][p] = c * a - sn * b; X[k][q] = sn * a + c * b; }
        }
    DMatrix A(d, std::vector<double>(d));
    for (int j = 0; j < n; j++) {
        ld nx = 0; for (int a = 0; a < n; a++) nx += X[a][j] * X[a][j];
        const ld l = sqrtl(S[j][j]), sc = sqrtl(1 / (2 * l) / nx);
        for (int k = 0; k < n; k++) { A[k][j] = (double)(sc * X[k][j]); A[n + k][j] = (double)(l * sc * X[k][j]); A[k][n + j] = A[k][j]; A[n + k][n + j] = -A[n + k][j]; }
    }
    localManifold base;
    base.dimension = d; base.stable = false; base.L = interval(0.05); base.params = par;
    base.F.A = A; base.F.p = zero;
    base.F.Ainv = inverseEnclosure(toInterval(A));
    {   // inverseEnclosure: A * Ainv contains the identity tightly
        const IMatrix P = matMul(toInterval(A), base.F.Ainv);
        for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) { const double e = i == j ? 1.0 : 0.0; CHECK(P[i][j].lo <= e && e <= P[i][j].hi && P[i][j].hi - P[i][j].lo < 1e-12); }
    }
    auto DF_over = [&](const localManifold& m, const IVector& U) {
        IVector x(d);
        for (int i = 0; i < d; i++) { interval a = m.F.p[i]; for (int j = 0; j < d; j++) a = ia::add(a, ia::mul(interval(m.F.A[i][j]), U[j])); x[i] = a; }
        localManifold_Eig t(m); t.F.p = x;                                    // Df at the box: reuse the mirror's own derivative formula
        return matMul(matMul(m.F.Ainv, t.Df_at_p()), toInterval(m.F.A));
    };
    double eps1 = 0;
    for (int pass = 0; pass < 2; pass++) {
        const double rad = pass == 0 ? 1e-3 : 2e-3;
        localManifold_Eig E(base);
        E.U_flat_global = IVector(n, interval(-rad, rad));
        const IVector U = E.constructU(E.U_flat_global);
        CHECK(E.checkConditionsOn(DF_over(E, U), U));                          // sets xi
        E.computeEigenError_minus_infty();
        // validated eigen-data
        for (int i = 0; i < d; i++) {
            CHECK(E.eigenvalues[i].hi - E.eigenvalues[i].lo < 1e-12);
            CHECK(i < n ? E.eigenvalues[i].lo > 0.6 : E.eigenvalues[i].hi < -0.6);
            for (int k = 0; k < d; k++) CHECK(E.Eigenvector_Error[k][i].hi - E.Eigenvector_Error[k][i].lo < 1e-12);
        }
        // K against the condition number in long double
        std::vector<std::vector<ld>> Al(d, std::vector<ld>(d));
        for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) Al[i][j] = A[i][j];
        const ld cond = two_norm_ld(Al) / two_norm_ld(Al, true);
        CHECK((ld)E.K_store.lo >= cond * (1 - 1e-9L) && (ld)E.K_store.hi <= cond * (1 + 1e-6L));
        // eps against the formula evaluated independently at the centre of the neighbourhood
        std::vector<ld> sl(n);
        for (int k = 0; k < n; k++) {
            std::vector<std::vector<ld>> S(n, std::vector<ld>(n, 0));
            const ld b[3] = {1.0L, 0.975L, 0.95L}, c[2] = {0.05L, -0.05L};
            for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
                ld v = 0;                                                    // Taylor coefficient of u_j u_k in f_{n+i} at u = 0
                if (i == j && j == k) v = -b[i] * 3 / 2;
                else if (j == k && std::abs(i - j) == 1) v = c[std::min(i, j)] * (0 - 0.5L);
                else if (j != k && (i == j || i == k) && std::abs(j - k) == 1) v = -c[std::min(j, k)];
                S[i][j] = v;
            }
            sl[k] = two_norm_ld(S);
        }
        const ld CG = sqrtl(sl[0] * sl[0] + sl[1] * sl[1] + sl[2] * sl[2]);
        const ld lam = cond * CG * sqrtl(3.0L) * rad * two_norm_ld(Al) * sqrtl(1 + 0.0025L) / (ld)E.xi.lo;
        const ld eps = lam / (1 - lam);
        CHECK((ld)E.eps_unscaled.hi >= eps && (ld)E.eps_unscaled.hi <= eps * 1.02L && E.eps_unscaled.lo > 0);
        CHECK(E.checkConjugatePointsBelowLminus());
        // Eu_m_Error_Final: symmetric boxes around zero of the size of eps in eigen-coordinates
        ccp_front_desc desc{};
        E.fillDescriptor(desc);
        double mx = 0;
        for (int i = 0; i < n; i++) for (int k = 0; k < n; k++) {
            const interval e = E.Eu_m_Error_Final[i][k];
            CHECK(e.lo < 0 && e.hi > 0 && std::fabs(e.lo + e.hi) < 1e-9 * e.hi);
            CHECK(e.hi > 0.3 * E.eps_unscaled.hi && e.hi < 30 * E.eps_unscaled.hi);
            CHECK(desc.Eu_err[i * CCP_MAX_N + k].lo == e.lo && desc.Eu_err[i * CCP_MAX_N + k].hi == e.hi);
            CHECK(E.getEigenError_minus_infty(k)[n + i].hi == e.hi && E.getEigenError_minus_infty(k)[i].hi == 0);
            mx = std::max(mx, e.hi);
        }
        E.computeEigenError_plus_infty();
        CHECK((int)E.Eigenfunction_Error_plus_infty.size() == d && E.Eigenfunction_Error_plus_infty[0][d - 1].hi >= E.eps_unscaled.hi);
        std::printf("radius %.0e: K = %.6f, eps = %.6e, max |Eu_err| = %.6e, xi = %.6f\n", rad, E.K_store.hi, E.eps_unscaled.hi, mx, E.xi.lo);
        if (pass == 0) eps1 = E.eps_unscaled.hi; else CHECK(E.eps_unscaled.hi > 1.95 * eps1 && E.eps_unscaled.hi < 2.2 * eps1);   // eps ~ r_u
    }
    {   // a neighbourhood for which (2.12) fails although lambda < 1
        localManifold_Eig E(base);
        E.U_flat_global = IVector(n, interval(-0.025, 0.025));
        const IVector U = E.constructU(E.U_flat_global);
        E.checkRateCondition(DF_over(E, U));
        E.ErrorEigenfunction();
        bool threw = false;                                                   // eps > 1: I + E_u has no certified inverse any more
        try { E.computeEigenError_minus_infty(); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw == (E.eps_unscaled.hi > 1.0));
        std::printf("radius 2.5e-2: eps = %.4f, below-L_- test %d\n", E.eps_unscaled.hi, (int)E.checkConjugatePointsBelowLminus());
        CHECK(E.eps_unscaled.hi > 0.47 ? !E.checkConjugatePointsBelowLminus() : E.checkConjugatePointsBelowLminus());
    }
    std::printf("localManifold_Eig ok\n");
    return 0;
}
