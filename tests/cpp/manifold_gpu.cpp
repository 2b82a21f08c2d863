// GPU exercise of the host mirrors around boundDFU: localManifold::checkConditions, localManifold_Eig::computeEigenError_*,
// construct_Manifold_at_LPlus and lastEuFrame (host/ccp_manifold_conditions.hpp, ccp_local_manifold_eig.hpp, ccp_l_plus.hpp) with the
// real ccp_bound_dfu_batch call.  Exit code 0 = ok, 77 = no GPU.  The n = 3 defaults; both equilibria (0 and 1) share the eigenbasis.
#include "../../computing-conjugate-points_b200/host/ccp_l_plus.hpp"
#include <cstdio>

using namespace ccp_host;
typedef long double ld;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
    if (ccp_device_count() <= 0) { std::printf("no GPU: compiled and linked only\n"); return 77; }
    if (ccp_init(0) != 0) { std::printf("init failed: %s\n", ccp_last_error()); return 2; }
    const int n = 3, d = 6;
    const IVector par = {interval(1.0), interval(0.975), interval(0.95), interval(0.05), interval(-0.05)};
    // eigenbasis of Df(0) = Df(1) = [[0, I], [M, 0]] in long double, symplectic normalisation |x|^2 = 1 / (2 lambda)
    const double Mh[3][3] = {{0.5, 0.025, 0.0}, {0.025, 0.4875, -0.025}, {0.0, -0.025, 0.475}};
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
    base.dimension = d; base.params = par; base.subdivisionNUM = 15;
    base.F.A = A; base.F.Ainv = inverseEnclosure(toInterval(A)); base.F.p = IVector(d, interval(0.0));

    // ---- unstable manifold at 0: boundDFU on the GPU against the enclosure over the whole box on the host
    localManifold_Eig Eu(base);
    Eu.stable = false; Eu.L = interval(0.05); Eu.U_flat_global = IVector(n, interval(-1e-4, 1e-4));
    {
        const IVector U = Eu.constructU(Eu.U_flat_global);
        const IMatrix Mg = boundDFU(n, par, Eu.F, U, false, 15);
        IVector x(d);
        for (int i = 0; i < d; i++) { interval a = Eu.F.p[i]; for (int j = 0; j < d; j++) a = ia::add(a, ia::mul(interval(A[i][j]), U[j])); x[i] = a; }
        localManifold_Eig t(Eu); t.F.p = x;
        const IMatrix Mw = matMul(matMul(Eu.F.Ainv, t.Df_at_p()), toInterval(A));
        for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) {
            CHECK(Mg[i][j].lo >= Mw[i][j].lo - 1e-12 && Mg[i][j].hi <= Mw[i][j].hi + 1e-12);       // hull over sub-boxes lies in the whole-box enclosure
            CHECK(Mg[i][j].lo <= Mg[i][j].hi);
        }
        CHECK(Eu.checkConditionsOn(Mw, U));
    }
    CHECK(Eu.checkConditions());
    std::printf("checkConditions (unstable, GPU boundDFU) ok: xi = %.6f, mu = %.6f\n", Eu.xi.lo, Eu.mu.hi);
    Eu.computeEigenError_minus_infty();
    CHECK(Eu.eps_unscaled.hi > 0 && Eu.eps_unscaled.hi < 1e-2 && Eu.checkConjugatePointsBelowLminus());
    ccp_front_desc desc{};
    Eu.fillDescriptor(desc);
    std::printf("computeEigenError_minus_infty ok: eps = %.4e, Eu_err(0,0) = [%.4e, %.4e]\n", Eu.eps_unscaled.hi, desc.Eu_err[0].lo, desc.Eu_err[0].hi);

    // ---- stable manifold at 1 around a point near the end of a front: construct_Manifold_at_LPlus and lastEuFrame
    LPlusStage st;
    st.dimension = d; st.manifold_subdivision = 15; st.stable = base;
    st.stable.stable = true;
    for (int i = 0; i < n; i++) st.stable.F.p[i] = interval(1.0);
    const double loc[6] = {1e-9, -2e-9, 1e-9, 1e-4, -5e-5, 3e-5};
    IVector endPoint(d);
    for (int i = 0; i < d; i++) { double s = 0; for (int k = 0; k < d; k++) s += A[i][k] * loc[k]; endPoint[i] = interval(s - 1e-15, s + 1e-15); }
    localManifold_Eig big = st.construct_Manifold_at_LPlus(endPoint);
    CHECK(big.stable && big.checkConditions());
    CHECK(big.L.hi >= 2e-5 && big.U_flat_global[0].hi >= 1e-4 && big.U_flat_global[0].lo < 0);
    big.computeEigenError_plus_infty();
    std::printf("construct_Manifold_at_LPlus ok: L = %.3e, xi = %.6f, eps = %.4e\n", big.L.hi, big.xi.lo, big.eps_unscaled.hi);
    {
        ccp_front_result r{};
        const double G[6][3] = {{1.0, 0.2, -0.1}, {0.1, -0.8, 0.3}, {0.0, 0.4, 0.5}, {1e-3, 2e-3, 0}, {-2e-3, 1e-3, 1e-3}, {0, 0, -0.5e-3}};
        for (int j = 0; j < n; j++) for (int i = 0; i < d; i++) {
            double s = 0; for (int k = 0; k < d; k++) s += A[i][k] * G[k][j];
            r.last_frame[i * (CCP_MAX_N + 2) + j].lo = s - 1e-12; r.last_frame[i * (CCP_MAX_N + 2) + j].hi = s + 1e-12;
            r.first_frame[i * (CCP_MAX_N + 1) + j].lo = A[i][j] - 1e-12; r.first_frame[i * (CCP_MAX_N + 1) + j].hi = A[i][j] + 1e-12;
        }
        for (int i = 0; i < d; i++) {                                         // phi' at the left end: a combination that keeps every frame of full rank
            const double s = A[i][0] + 0.5 * A[i][1] - 0.7 * A[i][2];
            r.first_frame[i * (CCP_MAX_N + 1) + n].lo = s - 1e-12; r.first_frame[i * (CCP_MAX_N + 1) + n].hi = s + 1e-12;
        }
        const topFrame fr(n, r);
        const bool L_PLUS = st.lastEuFrame(fr, endPoint);
        std::printf("lastEuFrame = %d\n", (int)L_PLUS);
        CHECK(L_PLUS);
        // the hook of the frameDet shim
        propagateManifold pm(desc);
        pm.L_plus = [&](const topFrame& f, const IVector& e) { return st.lastEuFrame(f, e); };
        (void)pm;
    }
    std::printf("manifold stages on the GPU ok\n");
    return 0;
}
