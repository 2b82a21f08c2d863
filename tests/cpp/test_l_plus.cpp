// Host-only test of host/ccp_l_plus.hpp (mirror of the L_+ stage, propagateManifold.cpp:192-558, and topFrame::checkFirstFrame).
// The GPU-backed pieces (construct_Manifold_at_LPlus / lastEuFrame call boundDFU) are not run here; the arithmetic of
// Proposition 2.9 / Section 3.4 and the projections are checked on constructed inputs with known answers.
#include "../../computing-conjugate-points_b200/host/ccp_l_plus.hpp"
#include <cstdio>

using namespace ccp_host;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
    const int n = 3, d = 6;
    // ---- detCofactor
    {
        const IMatrix M = {{interval(2.0), interval(0.0), interval(1.0)}, {interval(1.0), interval(3.0), interval(-1.0)}, {interval(0.0), interval(5.0), interval(4.0)}};
        const interval dt = detCofactor(M);                                   // 2(12+5) - 0 + 1(5-0) = 39
        CHECK(dt.lo <= 39 && 39 <= dt.hi && dt.hi - dt.lo < 1e-12);
    }
    // ---- compute_epsilon_beta: orthonormal thin Gamma -> mu* = 1, eps_beta = |Beta|_2
    IMatrix Gamma(n, IVector(n - 1, interval(0.0))), Beta(n, IVector(n - 1, interval(0.0)));
    Gamma[0][0] = interval(1.0); Gamma[1][1] = interval(1.0);
    Beta[0][0] = interval(3e-3); Beta[1][1] = interval(-4e-3); Beta[2][0] = interval(4e-3);           // columns (3,0,4)e-3, (0,-4,0)e-3 -> |B| = 5e-3
    {
        const interval e = compute_epsilon_beta(Gamma, Beta);
        CHECK(e.hi >= 5e-3 && e.hi < 5e-3 * (1 + 1e-9));        // norms are upper bounds (thin), like the .right() the reference takes
        IMatrix G2 = Gamma;                                                   // thick Gamma: mu* = 1 - 2 |Gd^T Gc| - |Gd^T Gd| < 1
        for (auto& r : G2) for (auto& x : r) x = ia::add(x, interval(-0.01, 0.01));
        const interval e2 = compute_epsilon_beta(G2, Beta);
        CHECK(e2.hi > e.hi * 1.02 && e2.hi < e.hi * 1.2);
        IMatrix G3 = Gamma; G3[1][1] = interval(0.0); G3[0][1] = interval(1.0);                       // rank one: mu* cannot be bounded above 0
        const interval e3 = compute_epsilon_beta(G3, Beta);
        CHECK(e3.hi < 0);
        CHECK(!checkL_plus_local(G3, Beta, interval(1e-8), interval(0.7), interval(0.67), d));
    }
    // ---- checkL_plus_local against the formulas of Proposition 2.9 in plain doubles
    {
        const double e0 = 1e-6, eb = 5e-3, nu1 = 0.7, nun = 0.67, nn = 3;
        const double s = std::sqrt(2 * nn * nu1), CP = 2 * s / (1 - e0 * s), CQ = std::sqrt(2 * nn / nun) + s + 2 * e0 * nn;
        const double CM1 = e0 * CQ * (2 + e0 * CQ), CM2 = e0 * CP * (2 + e0 * CP) + 2 * eb * (1 + e0 * CP) + eb * eb;
        const double tot = 2 * e0 * (CQ + CP) + e0 * e0 * (CQ * CQ + CP * CP) + eb * eb + 2 * eb * (1 + e0 * CP)
                         + (1 + e0 * CQ) * (1 + e0 * CQ) * CM1 / std::fabs(1 - CM1) + (1 + e0 * CP + eb) * (1 + e0 * CP + eb) * CM2 / std::fabs(1 - CM2);
        CHECK(tot < 1);
        CHECK(checkL_plus_local(Gamma, Beta, interval(e0), interval(nu1), interval(nun), d));
        // the acceptance threshold in eps_beta: scan Beta's scale and compare the decision with the double formula
        int flips = 0; bool prev = true;
        for (double sc = 1; sc < 200; sc *= 1.05) {
            IMatrix B = Beta; for (auto& r : B) for (auto& x : r) x = ia::mul(x, interval(sc));
            const double b = eb * sc, cm2 = e0 * CP * (2 + e0 * CP) + 2 * b * (1 + e0 * CP) + b * b;
            const double t = 2 * e0 * (CQ + CP) + e0 * e0 * (CQ * CQ + CP * CP) + b * b + 2 * b * (1 + e0 * CP)
                           + (1 + e0 * CQ) * (1 + e0 * CQ) * CM1 / std::fabs(1 - CM1) + (1 + e0 * CP + b) * (1 + e0 * CP + b) * cm2 / std::fabs(1 - cm2);
            const bool expect = cm2 < 1 && t < 1;
            const bool got = checkL_plus_local(Gamma, B, interval(e0), interval(nu1), interval(nun), d);
            if (std::fabs(t - 1) > 1e-6 && std::fabs(cm2 - 1) > 1e-6) CHECK(got == expect);
            if (got != prev) { flips++; prev = got; }
        }
        CHECK(flips == 1);                                                    // accepted for small eps_beta, rejected for large, one threshold
        CHECK(!checkL_plus_local(Gamma, Beta, interval(0.3), interval(nu1), interval(nun), d));       // eps_0 sqrt(2 n nu_1) >= ... Neumann test fails
    }
    std::printf("Proposition 2.9 arithmetic ok\n");

    // ---- projectionGammaBeta: last_Frame = A_s [G; B] with zero eigenfunction error gives back [G; B] / max|.|
    LPlusStage st;
    st.dimension = d;
    DMatrix A(d, std::vector<double>(d, 0.0));
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) A[i][j] = (i == j ? 2.0 : 0.0) + 0.1 * std::sin(1.0 + 3 * i + 7 * j);
    st.stable.dimension = d; st.stable.F.A = A; st.stable.F.p = IVector(d, interval(0.0));
    st.stable.F.Ainv = inverseEnclosure(toInterval(A));
    {
        double loc[6][3] = {{1.0, 0.2, -0.1}, {0.1, -0.8, 0.3}, {0.0, 0.4, 0.5}, {1e-3, 2e-3, 0}, {-2e-3, 1e-3, 1e-3}, {0, 0, -0.5e-3}};
        IMatrix last(d, IVector(n + 2, interval(0.0)));
        for (int j = 0; j < n; j++) for (int i = 0; i < d; i++) { double s = 0; for (int k = 0; k < d; k++) s += A[i][k] * loc[k][j]; last[i][j] = interval(s - 1e-13, s + 1e-13); }
        const IMatrix Zero(d, IVector(d, interval(0.0)));
        const IMatrix U = st.projectionGammaBeta(last, Zero);
        for (int j = 0; j < n; j++) {
            double mx = 0; for (int i = 0; i < d; i++) mx = std::max(mx, std::fabs(loc[i][j]));
            for (int i = 0; i < d; i++) { const double e = loc[i][j] / mx; CHECK(U[i][j].lo <= e + 1e-12 && U[i][j].hi >= e - 1e-12 && U[i][j].hi - U[i][j].lo < 1e-10); }
        }
        // an eigenfunction error widens the enclosure but keeps the error-free coordinates inside
        IMatrix Err(d, IVector(d, interval(-1e-4, 1e-4)));
        const IMatrix Ue = st.projectionGammaBeta(last, Err);
        double wmax = 0;
        for (int j = 0; j < n; j++) for (int i = 0; i < d; i++) {
            const double mid = 0.5 * (U[i][j].lo + U[i][j].hi);
            CHECK(Ue[i][j].lo < mid && mid < Ue[i][j].hi);
            wmax = std::max(wmax, Ue[i][j].hi - Ue[i][j].lo);
        }
        std::printf("widest Gamma/Beta entry with a 1e-4 eigenfunction error: %.3e\n", wmax);
        CHECK(wmax > 1e-4 && wmax < 2e-2);
    }
    std::printf("projectionGammaBeta ok\n");

    // ---- checkFirstFrame + checkL_plus on a synthetic result record
    {
        ccp_front_result r{};
        // first_frame: columns U_0, U_1, U_2, phi'  (2n x (n+1));  U_2 parallel to phi' -> frames that keep both are rank deficient
        const double cols[4][6] = {{1, 0, 0, 0.7, 0, 0}, {0, 1, 0, 0, 0.7, 0}, {0, 0, 1, 0, 0, 0.7}, {0, 0, 2, 0, 0, 1.4}};
        for (int i = 0; i < d; i++) for (int k = 0; k <= n; k++) { r.first_frame[i * (CCP_MAX_N + 1) + k].lo = cols[k][i] - 1e-12; r.first_frame[i * (CCP_MAX_N + 1) + k].hi = cols[k][i] + 1e-12; }
        const topFrame fr(n, r);
        CHECK(!checkFirstFrame(fr, 0) && !checkFirstFrame(fr, 1) && checkFirstFrame(fr, 2));
        IMatrix U_coord(d, IVector(n, interval(0.0)));
        for (int j = 0; j < n; j++) { U_coord[j][j] = interval(1.0); U_coord[n + (j + 1) % n][j] = interval(2e-3); }
        const IVector ev = {interval(0.72), interval(0.70), interval(0.67), interval(-0.67), interval(-0.70), interval(-0.72)};
        CHECK(st.checkL_plus(fr, U_coord, interval(1e-6), ev));              // only k = 2 has a full-rank first frame, and it passes
        ccp_front_result r2 = r;
        for (int i = 0; i < d; i++) { r2.first_frame[i * (CCP_MAX_N + 1) + n].lo = -1e-12; r2.first_frame[i * (CCP_MAX_N + 1) + n].hi = 1e-12; }      // phi' not separated from 0: no full-rank frame
        CHECK(!st.checkL_plus(topFrame(n, r2), U_coord, interval(1e-6), ev));
        IMatrix Ubad = U_coord; for (int j = 0; j < n; j++) Ubad[n + (j + 1) % n][j] = interval(0.9);
        CHECK(!st.checkL_plus(fr, Ubad, interval(1e-6), ev));
    }
    std::printf("checkFirstFrame / checkL_plus ok\n");
    return 0;
}
