// ccp_local_manifold_eig.hpp -- host mirror of localManifold_Eig (localManifold.cpp:426-681): the producer of the hot path's input
// `Eu_m_Error_Final` (= ccp_front_desc::Eu_err, read by construct_InitCondU, propagateManifold.cpp:6-64).
//
//   void     ErrorEigenfunction()                 localManifold.cpp:426-497   lambda_{T-} = K C_G r_u |A_0| sqrt(1+L^2) / xi,  eps = lambda/(1-lambda)
//   interval computeK()                           localManifold.cpp:500-535   validated eigen-data (eigenvalues.cpp), K = |Q| |Q^-1|
//   bool     checkConjugatePointsBelowLminus()    localManifold.cpp:540-571   eps < 1/sqrt(n (1 + |M_-|))   (Prop. 2.5 / (2.12))
//   void     computeEigenError_minus_infty()      localManifold.cpp:574-633   Eu_m_Error_Final = E_s (I + E_u)^-1
//   void     computeEigenError_plus_infty()       localManifold.cpp:635-670
//   IVector  getEigenError_minus_infty(column)    localManifold.cpp:673-681
//   tensorNorm, compressTensor, identityMat       utils.cpp:497-545
//
// Small, serial per front, host only.  CAPD pieces and their replacements, all rigorous on their own:
//   (*f)(x, Df, Hf) with setDegree(2)   -> closed-form second derivatives of the front field (frontHessian), stored like CAPD's IHessian
//                                          as Taylor coefficients D^a f / a!  [CAPD-recollection: Hf(i,j,j) = f_jj / 2, Hf(i,j,k) = f_jk]
//   gaussInverseMatrix / krawczykInverse -> inverseEnclosure (approximate inverse C, E = I - C Q, |E|_inf < 1  =>  Q^-1 = C + E C + E^2 (I-E)^-1 C)
//   gauss(A, b)                          -> Ainv * b with the enclosure of A^-1 the caller already holds (LocalField::Ainv)
//   EuclNorm                             -> the certified bounds of ccp_manifold_conditions.hpp
#pragma once
#include "ccp_manifold_conditions.hpp"

namespace ccp_host {

inline IMatrix identityMat(int n) {                                         // utils.cpp:497-503
    IMatrix I(n, IVector(n, interval(0.0)));
    for (int i = 0; i < n; i++) I[i][i] = interval(1.0);
    return I;
}
inline IMatrix matMul(const IMatrix& A, const IMatrix& B) {
    const int r = (int)A.size(), m = (int)B.size(), c = (int)B[0].size();
    IMatrix C(r, IVector(c));
    for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) { interval a(0.0); for (int k = 0; k < m; k++) a = ia::add(a, ia::mul(A[i][k], B[k][j])); C[i][j] = a; }
    return C;
}
inline IMatrix toInterval(const DMatrix& A) {                                // utils.cpp:12-23
    IMatrix M(A.size(), IVector(A[0].size()));
    for (size_t i = 0; i < A.size(); i++) for (size_t j = 0; j < A[0].size(); j++) M[i][j] = interval(A[i][j]);
    return M;
}

// enclosure of the inverse of every matrix in Q (throws when the approximate inverse does not contract):
//   C ~ mid(Q)^-1,  E = I - C Q,  |E|_inf < 1  =>  Q^-1 = (I - E)^-1 C = C + E C + E^2 (I - E)^-1 C;
// the first-order term is an interval matrix product (entrywise tight for a thick Q), the rest is bounded in the infinity norm.
inline IMatrix inverseEnclosure(const IMatrix& Q) {
    const int n = (int)Q.size();
    DMatrix mid(n, std::vector<double>(n));
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) mid[i][j] = ia::mid(Q[i][j]);
    const DMatrix C = approxInverse(mid);
    const IMatrix Ci = toInterval(C);
    IMatrix E = matMul(Ci, Q);
    double nE = 0;
    for (int i = 0; i < n; i++) {
        double s = 0;
        for (int j = 0; j < n; j++) { E[i][j] = ia::sub(interval(i == j ? 1.0 : 0.0), E[i][j]); s = ia::up(s + ia::mag(E[i][j])); }
        nE = std::max(nE, s);
    }
    if (!(nE < 1)) throw std::runtime_error("inverseEnclosure: approximate inverse does not contract");
    const IMatrix EC = matMul(E, Ci);
    const double x = ia::up(ia::up(nE * nE) / ia::dn(1 - nE));               // |E^2 (I - E)^-1|_inf <= nE^2 / (1 - nE)
    IMatrix out(n, IVector(n));
    for (int j = 0; j < n; j++) {
        double cmax = 0; for (int k = 0; k < n; k++) cmax = std::max(cmax, std::fabs(C[k][j]));
        const double r = ia::up(x * cmax);
        for (int i = 0; i < n; i++) out[i][j] = ia::add(ia::add(Ci[i][j], EC[i][j]), interval(-r, r));
    }
    return out;
}

// second-order Taylor coefficients of the velocity rows of the front field over x: H[i][j][k] ~ IHessian(n+i, j, k), j,k < n
inline std::vector<IMatrix> frontHessian(int n, const IVector& params, const IVector& x) {
    std::vector<IMatrix> H(n, IMatrix(n, IVector(n, interval(0.0))));
    for (int i = 0; i < n; i++) {
        // f_{n+i} = -b_i u_i (u_i - 1/2)(1 - u_i) - sum_{j ~ i} c_ij u_j (1 - u_j)(u_i - 1/2)
        H[i][i][i] = ia::mul(interval(0.5), ia::neg(ia::mul(params[i], ia::sub(interval(3.0), ia::mul(interval(6.0), x[i])))));
        for (int s = -1; s <= 1; s += 2) {
            const int j = i + s;
            if (j < 0 || j >= n) continue;
            const interval c = params[n + std::min(i, j)];
            const interval mixed = ia::neg(ia::mul(c, ia::sub(interval(1.0), ia::mul(interval(2.0), x[j]))));     // d^2/du_i du_j
            H[i][i][j] = mixed; H[i][j][i] = mixed;
            H[i][j][j] = ia::mul(interval(0.5), ia::mul(ia::mul(interval(2.0), c), ia::sub(x[i], interval(0.5))));  // (1/2) d^2/du_j^2
        }
    }
    return H;
}
// utils.cpp:505-531: Euclidean norm of the vector of spectral norms of the slices A_k[i][j] = T(i,j,k)
inline interval tensorNorm(const std::vector<IMatrix>& T, int n) {
    IVector sliced(n);
    for (int k = 0; k < n; k++) {
        IMatrix A(n, IVector(n));
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) A[i][j] = T[i][j][k];
        sliced[k] = euclNorm(A);
    }
    return euclNorm(sliced);
}

struct localManifold_Eig : localManifold {
    IVector eigenvalues;
    interval K_store, eps_unscaled;
    IMatrix Eigenvector_Error, Eu_m_Error_Final, Eigenfunction_Error_plus_infty;

    localManifold_Eig() {}
    explicit localManifold_Eig(const localManifold& lM) : localManifold(lM) {}

    IMatrix Df_at_p() const {                                                // (*(*pF).f)[(*pF).p]
        const int n = dimension / 2;
        IMatrix J(dimension, IVector(dimension, interval(0.0)));
        auto g = [&](int i) { return ia::mul(F.p[i], ia::sub(interval(1.0), F.p[i])); };
        auto w = [&](int i) { return ia::sub(F.p[i], interval(0.5)); };
        auto one_m2 = [&](int i) { return ia::sub(interval(1.0), ia::mul(interval(2.0), F.p[i])); };
        for (int i = 0; i < n; i++) {
            J[i][n + i] = interval(1.0);
            interval d = ia::neg(ia::mul(params[i], ia::sub(ia::mul(interval(3.0), g(i)), interval(0.5))));
            if (i > 0) { d = ia::sub(d, ia::mul(params[n + i - 1], g(i - 1))); J[n + i][i - 1] = ia::neg(ia::mul(params[n + i - 1], ia::mul(one_m2(i - 1), w(i)))); }
            if (i + 1 < n) { d = ia::sub(d, ia::mul(params[n + i], g(i + 1))); J[n + i][i + 1] = ia::neg(ia::mul(params[n + i], ia::mul(one_m2(i + 1), w(i)))); }
            J[n + i][i] = d;
        }
        return J;
    }

    // localManifold.cpp:500-535
    interval computeK() {
        const IMatrix A_u = toInterval(F.A), A_infty = Df_at_p();
        const IMatrix Lambda = matMul(matMul(F.Ainv, A_infty), A_u);
        eigenvalues = boundEigenvalues(Lambda);
        IVector Lambda_vec(dimension);
        for (int i = 0; i < dimension; i++) Lambda_vec[i] = Lambda[i][i];
        const std::vector<IMatrix> out = boundEigenvectors(A_infty, A_u, Lambda_vec);
        IMatrix Q = out[0];
        for (int i = 0; i < dimension; i++) for (int j = 0; j < dimension; j++) Q[i][j] = ia::add(out[0][i][j], out[1][i][j]);
        const interval K = ia::mul(euclNorm(Q), euclNorm(inverseEnclosure(Q)));
        Eigenvector_Error = out[1];
        K_store = K;
        return K;
    }

    // localManifold.cpp:426-497; needs xi from checkRateCondition
    void ErrorEigenfunction() {
        const int n = dimension / 2;
        const IVector U = constructU(U_flat_global);
        IVector x(dimension);                                                // p + pi_1 A U
        for (int i = 0; i < dimension; i++) {
            interval a = F.p[i];
            if (i < n) for (int j = 0; j < dimension; j++) a = ia::add(a, ia::mul(interval(F.A[i][j]), U[j]));
            x[i] = a;
        }
        const interval C_G(tensorNorm(frontHessian(n, params, x), n).hi);
        const interval K = computeK();
        const interval eta = xi;
        const interval norm_A0 = euclNorm(toInterval(F.A));
        IVector absU(n);
        for (int i = 0; i < n; i++) absU[i] = ia::abs(U_flat_global[i]);
        const interval r_u = euclNorm(absU);
        interval lambda = ia::div(ia::mul(ia::mul(ia::mul(ia::mul(K, C_G), r_u), norm_A0), ia::sqrt(ia::add(interval(1.0), ia::sqr(L)))), eta);
        lambda = interval(lambda.hi);
        // lambda/(1 - lambda) is the sum of the Neumann series only for lambda < 1; the reference evaluates the quotient regardless
        // (a negative eps would then pass the L_- test) -- here lambda >= 1 gives eps = +infinity and every later test fails.
        eps_unscaled = (lambda.hi < 1) ? ia::div(lambda, ia::sub(interval(1.0), lambda)) : interval(INFINITY);
    }

    // localManifold.cpp:540-571
    bool checkConjugatePointsBelowLminus() const {
        const IMatrix M_minus = blockDecompose(Df_at_p(), dimension)[2];
        const interval M_bound = euclNorm(M_minus);
        const interval RHS_2p12 = ia::inv(ia::sqrt(ia::mul(interval((double)(dimension / 2)), ia::add(interval(1.0), M_bound))));
        return eps_unscaled.hi < RHS_2p12.lo;
    }

    // localManifold.cpp:574-633
    void computeEigenError_minus_infty() {
        const int n = dimension / 2;
        ErrorEigenfunction();
        const interval eps_local = ia::mul(eps_unscaled, interval(-1.0, 1.0));
        IMatrix Error_mat(dimension, IVector(n));
        for (int j = 0; j < n; j++) for (int i = 0; i < dimension; i++) Error_mat[i][j] = ia::add(Eigenvector_Error[i][j], eps_local);
        const IMatrix sol = matMul(F.Ainv, Error_mat);                       // gauss(A_u, column) for every column
        IMatrix E_u(n, IVector(n)), E_s(n, IVector(n));
        for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) { E_u[i][j] = sol[i][j]; E_s[i][j] = sol[i + n][j]; }
        IMatrix IE = identityMat(n);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) IE[i][j] = ia::add(IE[i][j], E_u[i][j]);
        Eu_m_Error_Final = matMul(E_s, inverseEnclosure(IE));
    }

    // localManifold.cpp:635-670
    void computeEigenError_plus_infty() {
        ErrorEigenfunction();
        const interval eps_local = ia::mul(eps_unscaled, interval(-1.0, 1.0));
        IMatrix Error_mat(dimension, IVector(dimension));
        for (int j = 0; j < dimension; j++) for (int i = 0; i < dimension; i++) Error_mat[i][j] = ia::add(Eigenvector_Error[i][j], eps_local);
        Eigenfunction_Error_plus_infty = Error_mat;
    }

    // localManifold.cpp:673-681
    IVector getEigenError_minus_infty(int columnNumber) const {
        const int n = dimension / 2;
        IVector v(dimension, interval(0.0));
        for (int i = 0; i < n; i++) v[i + n] = Eu_m_Error_Final[i][columnNumber];
        return v;
    }

    // fills the field of the hot path's descriptor that construct_InitCondU reads from this object
    void fillDescriptor(ccp_front_desc& d) const {
        const int n = dimension / 2;
        for (int i = 0; i < n; i++) for (int k = 0; k < n; k++) d.Eu_err[i * CCP_MAX_N + k] = to_c(Eu_m_Error_Final[i][k]);
    }
};

}  // namespace ccp_host
