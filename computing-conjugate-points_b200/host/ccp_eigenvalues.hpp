// ccp_eigenvalues.hpp -- host mirror of the reference's eigenvalues.cpp on plain outward-rounded intervals (no CAPD).
//
//   IVector boundEigenvalues(IMatrix B)                                   eigenvalues.cpp:4-19     Gershgorin discs of a nearly diagonal B
//   IVector F_eigenvector (A, v, lambda)                                  eigenvalues.cpp:137-167  (A v - lambda v, |pi_1 v|^2 - 1/(2|lambda|))
//   IMatrix DF_eigenvector(A, v, lambda)                                  eigenvalues.cpp:169-205  its derivative in (v, lambda)
//   vector<IVector> krawczykEigenvector(A, V, lambda, H_vec)              eigenvalues.cpp:90-135   one Krawczyk step, returns {image, new_Nbd}
//   vector<IVector> boundSingleEigenvector(A, V, lambda)                  eigenvalues.cpp:21-88    5 steps from a 3e-14 box, throws -344
//   vector<IMatrix> boundEigenvectors(A, Q, Lambda)                       eigenvalues.cpp:207-256  column by column, returns {Q_center, Q_error}
//
// These feed localManifold_Eig::computeK (localManifold.cpp:499-535): `Eigenvector_Error` = Q_error is what, scaled, becomes the
// descriptor field `Eu_err` of the hot path (localManifold.cpp:673-681).  SURVEY.md 8(f) rank 4; small, serial per front, host only.
// Same names, argument meaning and error behaviour as the reference; matrices are row-major vectors of rows.
#pragma once
#include "ccp_bvp.hpp"

namespace ccp_host {

namespace ia {
inline interval abs(const interval& a) {
    if (a.lo >= 0) return a;
    if (a.hi <= 0) return interval(-a.hi, -a.lo);
    return interval(0.0, std::max(-a.lo, a.hi));
}
// 1/a for an interval that does not contain zero (throws otherwise, like CAPD's division)
inline interval inv(const interval& a) {
    if (!(a.lo > 0 || a.hi < 0)) throw std::domain_error("interval division by an interval containing zero");
    return interval(dn(1.0 / a.hi), up(1.0 / a.lo));
}
inline bool subsetInterior(const interval& a, const interval& b) { return b.lo < a.lo && a.hi < b.hi; }
}

// point approximation of the inverse of a point matrix (Gauss-Jordan, partial pivoting); the reference takes
// midMatrix(gaussInverseMatrix(midMatrix(.))) -- a Krawczyk operator is valid with ANY approximate inverse
inline DMatrix approxInverse(const DMatrix& Min) {
    const int D = (int)Min.size();
    DMatrix M(D, std::vector<double>(2 * D, 0.0));
    for (int i = 0; i < D; i++) { for (int j = 0; j < D; j++) M[i][j] = Min[i][j]; M[i][D + i] = 1.0; }
    for (int c = 0; c < D; c++) {
        int piv = c;
        for (int r = c + 1; r < D; r++) if (std::fabs(M[r][c]) > std::fabs(M[piv][c])) piv = r;
        std::swap(M[c], M[piv]);
        const double dv = M[c][c];
        if (dv == 0.0 || !std::isfinite(dv)) throw std::runtime_error("approxInverse: singular matrix");
        for (int j = 0; j < 2 * D; j++) M[c][j] /= dv;
        for (int r = 0; r < D; r++) if (r != c) { const double f = M[r][c]; for (int j = 0; j < 2 * D; j++) M[r][j] -= f * M[c][j]; }
    }
    DMatrix out(D, std::vector<double>(D));
    for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) out[i][j] = M[i][D + j];
    return out;
}

// eigenvalues.cpp:4-19
inline IVector boundEigenvalues(const IMatrix& B) {
    const int k = (int)B.size();
    IVector ev(k);
    for (int i = 0; i < k; i++) {
        interval sum(0.0);
        for (int j = 0; j < k; j++) if (i != j) sum = ia::add(sum, ia::abs(B[i][j]));
        ev[i] = ia::add(B[i][i], ia::mul(interval(-1.0, 1.0), sum));
    }
    return ev;
}

// eigenvalues.cpp:137-167
inline IVector F_eigenvector(const IMatrix& A, const IVector& v, const interval& lambda) {
    const int d = (int)A.size();
    IVector F(d + 1);
    for (int i = 0; i < d; i++) {
        interval a(0.0);
        for (int j = 0; j < d; j++) a = ia::add(a, ia::mul(A[i][j], v[j]));
        F[i] = ia::sub(a, ia::mul(lambda, v[i]));
    }
    interval nrm(0.0);
    for (int i = 0; i < d / 2; i++) nrm = ia::add(nrm, ia::sqr(v[i]));                 // |pi_1 v|^2, normalisation (2.18)/(2.19)
    F[d] = ia::sub(nrm, ia::inv(ia::mul(interval(2.0), ia::abs(lambda))));
    return F;
}

// eigenvalues.cpp:169-205
inline IMatrix DF_eigenvector(const IMatrix& A, const IVector& v, const interval& lambda) {
    const int d = (int)A.size();
    IMatrix DF(d + 1, IVector(d + 1, interval(0.0)));
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) DF[i][j] = (i == j) ? ia::sub(A[i][j], lambda) : A[i][j];
    for (int i = 0; i < d; i++) DF[i][d] = ia::neg(v[i]);
    for (int j = 0; j < d / 2; j++) DF[d][j] = ia::mul(interval(2.0), v[j]);
    const interval q = ia::inv(ia::mul(interval(2.0), ia::sqr(lambda)));               // d/dlambda of -1/(2|lambda|) = sgn(lambda)/(2 lambda^2)
    DF[d][d] = (lambda.lo > 0) ? q : ia::neg(q);
    return DF;
}

// eigenvalues.cpp:90-135:  new_Nbd = -C F(z) + (I - C DF([z])) ([z] - z),  image = z + new_Nbd
inline std::vector<IVector> krawczykEigenvector(const IMatrix& A, const IVector& V, const interval& lambda, const IVector& H_vec) {
    const int d = (int)A.size();
    const interval lambda_nbd = ia::add(lambda, H_vec[d]);
    IVector V_nbd(d);
    for (int i = 0; i < d; i++) V_nbd[i] = ia::add(V[i], H_vec[i]);
    const IVector F = F_eigenvector(A, V, lambda);
    const IMatrix DF = DF_eigenvector(A, V, lambda), DFH = DF_eigenvector(A, V_nbd, lambda_nbd);
    DMatrix mid(d + 1, std::vector<double>(d + 1));
    for (int i = 0; i <= d; i++) for (int j = 0; j <= d; j++) mid[i][j] = 0.5 * (DF[i][j].lo + DF[i][j].hi);
    const DMatrix C = approxInverse(mid);
    IVector new_Nbd(d + 1), image(d + 1);
    for (int i = 0; i <= d; i++) {
        interval a(0.0);
        for (int k = 0; k <= d; k++) a = ia::sub(a, ia::mul(interval(C[i][k]), F[k]));
        for (int j = 0; j <= d; j++) {
            interval e(i == j ? 1.0 : 0.0);
            for (int k = 0; k <= d; k++) e = ia::sub(e, ia::mul(interval(C[i][k]), DFH[k][j]));
            a = ia::add(a, ia::mul(e, H_vec[j]));
        }
        new_Nbd[i] = a;
        image[i] = ia::add(i < d ? V[i] : lambda, a);
    }
    return {image, new_Nbd};
}

// eigenvalues.cpp:21-88: five Krawczyk steps from the box 3e-14 [-1,1]^(d+1), each inflated by (1 + 1e-10); the eigenvector components
// of the last step must map into the interior of their box, otherwise `throw -344`.  Returns {V, V_err} (V is not updated).
inline std::vector<IVector> boundSingleEigenvector(const IMatrix& A, const IVector& V, const interval& lambda) {
    const int d = (int)A.size();
    const int Newton_Steps = 5;
    const double scale = 3e-14;
    IVector H_vec(d + 1, interval(-scale, scale));
    bool verify = true;
    for (int it = 0; it < Newton_Steps; it++) {
        const std::vector<IVector> out = krawczykEigenvector(A, V, lambda, H_vec);
        const IVector& new_Nbd = out[1];
        verify = true;
        for (int i = 0; i < d; i++) if (!ia::subsetInterior(new_Nbd[i], H_vec[i])) verify = false;
        if (it == Newton_Steps - 1) break;
        for (int i = 0; i <= d; i++) H_vec[i] = ia::mul(new_Nbd[i], interval(1.0 + 1e-10));
    }
    if (!verify) throw -344;
    IVector V_err(H_vec.begin(), H_vec.begin() + d);
    return {V, V_err};
}

// eigenvalues.cpp:207-256: columns of Q are approximate eigenvectors of A for the approximate eigenvalues Lambda
inline std::vector<IMatrix> boundEigenvectors(const IMatrix& A, const IMatrix& Q, const IVector& Lambda) {
    const int d = (int)A.size();
    IMatrix Q_center(d, IVector(d)), Q_error(d, IVector(d));
    for (int j = 0; j < d; j++) {
        IVector col(d);
        for (int i = 0; i < d; i++) col[i] = Q[i][j];
        const std::vector<IVector> out = boundSingleEigenvector(A, col, Lambda[j]);
        for (int i = 0; i < d; i++) { Q_center[i][j] = out[0][i]; Q_error[i][j] = out[1][i]; }
    }
    return {Q_center, Q_error};
}

}  // namespace ccp_host
