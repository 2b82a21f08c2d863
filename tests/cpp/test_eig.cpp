// Host-only test of host/ccp_eigenvalues.hpp (mirror of the reference's eigenvalues.cpp).  No GPU, no oracle: the checks are
// mathematical -- the validated boxes must contain eigenpairs refined independently in long double (64-bit mantissa).
#include "../../computing-conjugate-points_b200/host/ccp_eigenvalues.hpp"
#include <cstdio>
#include <cstdlib>

using namespace ccp_host;
typedef long double ld;

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

// cyclic Jacobi on a symmetric n x n matrix in long double: eigenvalues on the diagonal, eigenvectors in the columns of X
static void jacobi(int n, std::vector<std::vector<ld>>& S, std::vector<std::vector<ld>>& X) {
    X.assign(n, std::vector<ld>(n, 0)); for (int i = 0; i < n; i++) X[i][i] = 1;
    for (int sweep = 0; sweep < 30; sweep++)
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
            if (S[p][q] == 0) continue;
            const ld th = (S[q][q] - S[p][p]) / (2 * S[p][q]);
            const ld t = (th >= 0 ? 1 : -1) / (fabsl(th) + sqrtl(th * th + 1)), c = 1 / sqrtl(t * t + 1), s = t * c;
            for (int k = 0; k < n; k++) { const ld a = S[k][p], b = S[k][q]; S[k][p] = c * a - s * b; S[k][q] = s * a + c * b; }
            for (int k = 0; k < n; k++) { const ld a = S[p][k], b = S[q][k]; S[p][k] = c * a - s * b; S[q][k] = s * a + c * b; }
            for (int k = 0; k < n; k++) { const ld a = X[k][p], b = X[k][q]; X[k][p] = c * a - s * b; X[k][q] = s * a + c * b; }
        }
}

int main() {
    const int n = 3, d = 2 * n;
    // M = R diag(mu) R^T evaluated in double; from here on the DOUBLE matrix is the object whose eigenpairs are enclosed
    const double mu[3] = {0.5, 1.3, 2.1}, a1 = 0.3, a2 = -0.7, a3 = 1.1;
    double R[3][3];
    {
        const double c1 = std::cos(a1), s1 = std::sin(a1), c2 = std::cos(a2), s2 = std::sin(a2), c3 = std::cos(a3), s3 = std::sin(a3);
        const double Rz[3][3] = {{c1, -s1, 0}, {s1, c1, 0}, {0, 0, 1}}, Ry[3][3] = {{c2, 0, s2}, {0, 1, 0}, {-s2, 0, c2}}, Rx[3][3] = {{1, 0, 0}, {0, c3, -s3}, {0, s3, c3}};
        double T[3][3];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { T[i][j] = 0; for (int k = 0; k < 3; k++) T[i][j] += Rz[i][k] * Ry[k][j]; }
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { R[i][j] = 0; for (int k = 0; k < 3; k++) R[i][j] += T[i][k] * Rx[k][j]; }
    }
    double M[3][3];
    for (int i = 0; i < 3; i++) for (int j = i; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += R[i][k] * mu[k] * R[j][k]; M[i][j] = M[j][i] = s; }
    IMatrix A(d, IVector(d, interval(0.0)));
    for (int i = 0; i < n; i++) { A[i][n + i] = interval(1.0); for (int j = 0; j < n; j++) A[n + i][j] = interval(M[i][j]); }

    // independent eigenpairs of A = [[0,I],[M,0]]: lambda = +-sqrt(mu_i), v = (x, lambda x), |x|^2 = 1/(2|lambda|)
    std::vector<std::vector<ld>> S(n, std::vector<ld>(n)), X;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) S[i][j] = M[i][j];
    jacobi(n, S, X);
    std::vector<ld> lam(d); std::vector<std::vector<ld>> V(d, std::vector<ld>(d));
    for (int j = 0; j < d; j++) {
        const int i = j % n; const ld l = (j < n ? 1 : -1) * sqrtl(S[i][i]);
        ld nx = 0; for (int k = 0; k < n; k++) nx += X[k][i] * X[k][i];
        const ld sc = sqrtl(1 / (2 * fabsl(l)) / nx);
        lam[j] = l;
        for (int k = 0; k < n; k++) { V[k][j] = sc * X[k][i]; V[n + k][j] = l * sc * X[k][i]; }
    }
    IMatrix Q(d, IVector(d)); IVector Lambda(d);
    for (int j = 0; j < d; j++) { Lambda[j] = interval((double)lam[j]); for (int i = 0; i < d; i++) Q[i][j] = interval((double)V[i][j]); }

    // boundEigenvectors: every column validated, boxes tiny, and they contain the long double eigenvectors
    const std::vector<IMatrix> out = boundEigenvectors(A, Q, Lambda);
    double worst = 0;
    for (int j = 0; j < d; j++) for (int i = 0; i < d; i++) {
        const interval c = out[0][i][j], e = out[1][i][j];
        CHECK(c.lo == Q[i][j].lo && c.hi == Q[i][j].hi);                       // centre is returned unchanged
        CHECK(e.lo < 0 && e.hi > 0 && e.hi - e.lo < 1e-12);
        const ld dev = V[i][j] - (ld)c.lo;                                      // true - approximate must lie in the error box
        CHECK(dev >= (ld)e.lo && dev <= (ld)e.hi);
        worst = std::max(worst, e.hi - e.lo);
    }
    std::printf("boundEigenvectors ok: %d columns validated, widest error box %.3e\n", d, worst);

    // one Krawczyk step by hand: the image of the box lies in its interior and contains the true eigenvalue
    {
        IVector col(d); for (int i = 0; i < d; i++) col[i] = Q[i][0];
        const IVector H(d + 1, interval(-3e-14, 3e-14));
        const std::vector<IVector> k = krawczykEigenvector(A, col, Lambda[0], H);
        for (int i = 0; i <= d; i++) CHECK(ia::subsetInterior(k[1][i], H[i]));
        CHECK((ld)k[0][d].lo <= lam[0] && lam[0] <= (ld)k[0][d].hi);
        const IVector F = F_eigenvector(A, col, Lambda[0]);
        for (int i = 0; i <= d; i++) CHECK(F[i].lo <= 1e-15 && F[i].hi >= -1e-15 && F[i].hi - F[i].lo < 1e-14);
        const IMatrix DF = DF_eigenvector(A, col, Lambda[0]), DFn = DF_eigenvector(A, col, Lambda[n]);
        CHECK(DF[d][d].lo > 0 && DFn[d][d].hi < 0);                           // sgn(lambda) / (2 lambda^2)
        CHECK(DF[0][d].lo == -col[0].hi && DF[d][0].lo <= 2 * col[0].lo && DF[d][n].lo == 0 && DF[d][n].hi == 0);
        std::printf("krawczykEigenvector ok\n");
    }

    // failure behaviour: an eigenvector that is off by more than the 3e-14 start box cannot be validated -> throw -344
    {
        IVector col(d); for (int i = 0; i < d; i++) col[i] = Q[i][1];
        col[0] = interval(col[0].lo + 1e-9);
        int code = 0;
        try { boundSingleEigenvector(A, col, Lambda[1]); } catch (int e) { code = e; }
        CHECK(code == -344);
        bool dom = false;
        try { F_eigenvector(A, col, interval(-1e-3, 1e-3)); } catch (const std::domain_error&) { dom = true; }
        CHECK(dom);
        std::printf("failure codes ok\n");
    }

    // boundEigenvalues: Gershgorin discs of Lambda = Q^-1 A Q (approximately diagonal) contain the true eigenvalues
    {
        DMatrix Qm(d, std::vector<double>(d));
        for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) Qm[i][j] = Q[i][j].lo;
        const DMatrix Qi = approxInverse(Qm);
        IMatrix B(d, IVector(d));
        for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) {
            interval s(0.0);
            for (int k = 0; k < d; k++) for (int l = 0; l < d; l++) s = ia::add(s, ia::mul(ia::mul(interval(Qi[i][k]), A[k][l]), Q[l][j]));
            B[i][j] = s;
        }
        const IVector ev = boundEigenvalues(B);
        for (int i = 0; i < d; i++) {
            double off = 0; for (int j = 0; j < d; j++) if (j != i) off += std::max(std::fabs(B[i][j].lo), std::fabs(B[i][j].hi));
            CHECK(ev[i].lo <= B[i][i].lo - off && ev[i].hi >= B[i][i].hi + off);
            CHECK(ev[i].hi - ev[i].lo < 1e-13);
            CHECK(std::fabs((double)lam[i] - 0.5 * (ev[i].lo + ev[i].hi)) < 1e-13);     // Q^-1 is approximate: discs are near, not around
        }
        const IMatrix Z = {{interval(2.0), interval(-0.5, 0.25)}, {interval(0.125), interval(-1.0)}};
        const IVector z = boundEigenvalues(Z);
        CHECK(z[0].lo <= 1.5 && z[0].lo > 1.5 - 1e-15 && z[0].hi >= 2.5 && z[0].hi < 2.5 + 1e-15 && z[1].lo <= -1.125 && z[1].hi >= -0.875);
        std::printf("boundEigenvalues ok\n");
    }
    return 0;
}
