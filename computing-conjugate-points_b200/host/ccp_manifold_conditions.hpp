// ccp_manifold_conditions.hpp -- host mirror of the caller of boundDFU: localManifold::checkConditions and what it needs.
//
//   IVector localManifold::constructU(IVector U_flat)                     localManifold.cpp:4-41
//   bool    localManifold::checkIsolatingBlock(IMatrix DFU, IVector U)    localManifold.cpp:180-239
//   bool    localManifold::checkRateCondition(IMatrix DFU)                localManifold.cpp:242-304   (sets the class variable xi)
//   bool    localManifold::checkConditions()                              localManifold.cpp:307-325   = constructU -> boundDFU -> both checks
//   blockDecompose, getMax, ml                                            utils.cpp:343-393, 448-452
//   localVField::image(0) = Ainv * f(p)                                   localVField.h:32
//
// boundDFU is the GPU call (ccp_local_manifold.hpp); everything else here is a handful of n x n interval operations per front and stays
// on the host.  CAPD's norm classes are replaced by bounds of the same meaning that are rigorous on their own:
//   EuclNorm(vector)  = sqrt(sum x_i^2), upper end
//   EuclNorm(matrix)  = sqrt(lambda_max(A^T A)),  EuclLNorm(matrix) = lambda_max((A + A^T)/2)      [CAPD-recollection: CAPD bounds both by
//   diagonalising the symmetric matrix approximately and applying Gershgorin].  Here lambda_max(S) <= mu is certified by congruence:
//   with X the floating-point eigenvectors of mid(S), mu D - C (C = X^T S X, D = X^T X, both in interval arithmetic) is diagonally
//   dominant with a non-negative diagonal  =>  positive semidefinite  =>  mu I - S is (Sylvester), for EVERY symmetric matrix in S.
//   X need not be orthogonal for this, so no rounding error of the eigen-solver enters the argument.
#pragma once
#include "ccp_eigenvalues.hpp"
#include "ccp_local_manifold.hpp"

namespace ccp_host {

namespace ia {
inline double mag(const interval& a) { return std::max(std::fabs(a.lo), std::fabs(a.hi)); }
inline double mid(const interval& a) { return 0.5 * (a.lo + a.hi); }
inline interval div(const interval& a, const interval& b) { return mul(a, inv(b)); }
}

// ---- symmetric eigen-solver in floating point (cyclic Jacobi); only an approximation is needed
inline void jacobiEigenvectors(DMatrix S, DMatrix& X) {
    const int n = (int)S.size();
    X.assign(n, std::vector<double>(n, 0.0));
    for (int i = 0; i < n; i++) X[i][i] = 1.0;
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0;
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += S[p][q] * S[p][q];
        if (off == 0.0) break;
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
            if (S[p][q] == 0.0) continue;
            const double th = (S[q][q] - S[p][p]) / (2 * S[p][q]);
            const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1)), c = 1 / std::sqrt(t * t + 1), s = t * c;
            for (int k = 0; k < n; k++) { const double a = S[k][p], b = S[k][q]; S[k][p] = c * a - s * b; S[k][q] = s * a + c * b; }
            for (int k = 0; k < n; k++) { const double a = S[p][k], b = S[q][k]; S[p][k] = c * a - s * b; S[q][k] = s * a + c * b; }
            for (int k = 0; k < n; k++) { const double a = X[k][p], b = X[k][q]; X[k][p] = c * a - s * b; X[k][q] = s * a + c * b; }
        }
    }
}

// rigorous upper bound of the largest eigenvalue of every symmetric matrix contained in S (S is symmetrised first)
inline double maxEigenvalueSym(const IMatrix& Sin) {
    const int n = (int)Sin.size();
    IMatrix S(n, IVector(n));
    DMatrix M(n, std::vector<double>(n));
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        S[i][j] = ia::mul(interval(0.5), ia::add(Sin[i][j], Sin[j][i]));
        M[i][j] = 0.5 * (ia::mid(Sin[i][j]) + ia::mid(Sin[j][i]));
    }
    DMatrix X;
    jacobiEigenvectors(M, X);
    IMatrix C(n, IVector(n)), D(n, IVector(n));
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        interval c(0.0), d(0.0);
        for (int k = 0; k < n; k++) {
            interval sx(0.0);
            for (int l = 0; l < n; l++) sx = ia::add(sx, ia::mul(S[k][l], interval(X[l][j])));
            c = ia::add(c, ia::mul(interval(X[k][i]), sx));
            d = ia::add(d, ia::mul(interval(X[k][i]), interval(X[k][j])));
        }
        C[i][j] = c; D[i][j] = d;
    }
    double mu = -INFINITY;
    for (int i = 0; i < n; i++) {
        double num = C[i][i].hi, offD = 0;
        for (int j = 0; j < n; j++) if (j != i) { num = ia::up(num + ia::mag(C[i][j])); offD = ia::up(offD + ia::mag(D[i][j])); }
        const double den = ia::dn(D[i][i].lo - offD);
        // non-finite input (NaN / inf entries) must never certify a bound: NaN compares false everywhere, which would otherwise
        // leave mu at -inf (euclNorm 0, ml +inf).  +inf is the only sound answer.
        if (!std::isfinite(num) || !std::isfinite(den) || !std::isfinite(offD) || !std::isfinite(D[i][i].hi)) return INFINITY;
        if (!(den > 0)) return INFINITY;               // D = X^T X strictly diagonally dominant  =>  X invertible (needed for Sylvester)
        // row i of mu D - C is dominant once  mu (D_ii - offD) >= num  for mu >= 0  (any mu >= 0 if num < 0),
        //                                     mu (D_ii + offD) >= num  for mu <  0  (needs num < 0)
        const double mi = (num >= 0) ? ia::up(num / den) : ia::up(num / ia::up(D[i][i].hi + offD));
        if (!std::isfinite(mi)) return INFINITY;
        mu = std::max(mu, mi);
    }
    return mu;
}

inline interval euclNorm(const IVector& v) {                                 // upper bound, returned as a thin interval like CAPD's right()
    interval s(0.0);
    for (const interval& x : v) s = ia::add(s, ia::sqr(x));
    return interval(ia::sqrt(s).hi);
}
inline interval euclNorm(const IMatrix& A) {                                 // sqrt(lambda_max(A^T A))
    const int r = (int)A.size(), c = (int)A[0].size();
    IMatrix ATA(c, IVector(c));
    for (int i = 0; i < c; i++) for (int j = 0; j < c; j++) {
        interval s(0.0);
        for (int k = 0; k < r; k++) s = ia::add(s, (i == j) ? ia::sqr(A[k][i]) : ia::mul(A[k][i], A[k][j]));
        ATA[i][j] = s;
    }
    const double m = std::max(0.0, maxEigenvalueSym(ATA));
    return interval(ia::up(std::sqrt(m)));
}
inline interval euclLNorm(const IMatrix& A) { return interval(maxEigenvalueSym(A)); }           // logarithmic norm, upper bound
inline interval ml(const IMatrix& A) {                                                          // utils.cpp:448-452:  -l(-A), a LOWER bound
    IMatrix N = A;
    for (auto& row : N) for (auto& x : row) x = ia::neg(x);
    return interval(-maxEigenvalueSym(N));
}

// utils.cpp:343-378
inline std::vector<IMatrix> blockDecompose(const IMatrix& M, int dimension) {
    const int n = dimension / 2;
    std::vector<IMatrix> out(4, IMatrix(n, IVector(n)));
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        out[0][i][j] = M[i][j]; out[1][i][j] = M[i][j + n]; out[2][i][j] = M[i + n][j]; out[3][i][j] = M[i + n][j + n];
    }
    return out;
}
// utils.cpp:380-391: the largest right end
inline interval getMax(const IVector& V) {
    double m = V[0].hi;
    for (size_t i = 1; i < V.size(); i++) if (V[i].hi > m) m = V[i].hi;
    return interval(m);
}

// the front ODE's field, written as the reference's formula string (ODE_functions.cpp:222): n-chain, parameters (b_1..b_n, c_12..c_{n-1,n})
inline IVector frontField(int n, const IVector& params, const IVector& x) {
    IVector f(2 * n);
    auto g = [&](int i) { return ia::mul(x[i], ia::sub(interval(1.0), x[i])); };        // u(1-u)
    auto w = [&](int i) { return ia::sub(x[i], interval(0.5)); };
    for (int i = 0; i < n; i++) {
        f[i] = x[n + i];
        interval a = ia::neg(ia::mul(params[i], ia::mul(ia::mul(x[i], w(i)), ia::sub(interval(1.0), x[i]))));
        if (i > 0) a = ia::sub(a, ia::mul(params[n + i - 1], ia::mul(g(i - 1), w(i))));
        if (i + 1 < n) a = ia::sub(a, ia::mul(params[n + i], ia::mul(g(i + 1), w(i))));
        f[n + i] = a;
    }
    return f;
}

struct localManifold {
    int dimension = 6;               // 2n
    bool stable = false;
    interval L;                      // cone slope
    IVector U_flat_global;           // n local coordinates
    int subdivisionNUM = 15;
    IVector params;                  // b_1..b_n, c_12..
    LocalField F;                    // (*pF): A, Ainv, p
    interval xi, mu;                 // set by checkRateCondition (xi is read by ErrorEigenfunction, localManifold.cpp:478)

    // localManifold.cpp:4-41
    IVector constructU(const IVector& U_flat) const {
        const int n = dimension / 2;
        interval norm(0.0);
        for (int i = 0; i < n; i++) norm = ia::add(norm, ia::sqr(U_flat[i]));
        norm = ia::sqrt(norm);
        const interval error = ia::mul(ia::mul(L, interval(-1.0, 1.0)), norm);
        IVector U(dimension);
        for (int i = 0; i < dimension; i++) {
            if (stable) U[i] = (i < n) ? error : U_flat[i - n];
            else U[i] = (i < n) ? U_flat[i] : error;
        }
        return U;
    }
    // (*pF)(0) = Ainv * f(p)
    IVector F_of_zero() const {
        const IVector fp = frontField(dimension / 2, params, F.p);
        IVector out(dimension);
        for (int i = 0; i < dimension; i++) {
            interval a(0.0);
            for (int j = 0; j < dimension; j++) a = ia::add(a, ia::mul(F.Ainv[i][j], fp[j]));
            out[i] = a;
        }
        return out;
    }
    // localManifold.cpp:180-239
    bool checkIsolatingBlock(const IMatrix& DFU, const IVector& U) const {
        const int n = dimension / 2;
        const interval r1(ia::abs(U[0]).hi), r2(ia::abs(U[n]).hi);
        const std::vector<IMatrix> ABCD = blockDecompose(DFU, dimension);
        IMatrix R11 = ABCD[0], R22 = ABCD[3];
        const IMatrix &R12 = ABCD[1], &R21 = ABCD[2];
        IVector lambda(n), beta(n), mlambda(n);
        for (int i = 0; i < n; i++) {
            lambda[i] = interval(ia::mid(ABCD[0][i][i])); beta[i] = interval(ia::mid(ABCD[3][i][i]));
            mlambda[i] = ia::neg(lambda[i]);
            R11[i][i] = ia::sub(R11[i][i], lambda[i]); R22[i][i] = ia::sub(R22[i][i], beta[i]);
        }
        const interval lambda_min = ia::neg(getMax(mlambda)), beta_max = getMax(beta);
        const interval nF = euclNorm(F_of_zero());
        const interval unstable_bound = ia::sub(ia::sub(ia::sub(lambda_min, ia::mul(euclNorm(R11), r1)), nF), ia::mul(euclNorm(R12), r2));
        const interval stable_bound = ia::add(ia::add(ia::add(beta_max, ia::mul(euclNorm(R21), r1)), nF), ia::mul(euclNorm(R22), r2));
        return unstable_bound.lo > 0 && stable_bound.hi < 0;
    }
    // localManifold.cpp:242-304
    bool checkRateCondition(const IMatrix& DFU) {
        const std::vector<IMatrix> ABCD = blockDecompose(DFU, dimension);
        auto negm = [](IMatrix M) { for (auto& r : M) for (auto& x : r) x = ia::neg(x); return M; };
        const IMatrix Fxx = stable ? negm(ABCD[3]) : ABCD[0], Fxy = stable ? negm(ABCD[2]) : ABCD[1];
        const IMatrix Fyx = stable ? negm(ABCD[1]) : ABCD[2], Fyy = stable ? negm(ABCD[0]) : ABCD[3];
        mu = ia::add(euclLNorm(Fyy), ia::div(euclNorm(Fyx), L));
        xi = ia::sub(ml(Fxx), ia::mul(L, euclNorm(Fxy)));
        return mu.hi < 0 && 0 < xi.lo;
    }
    // the two checks on an enclosure M of DF over U (localManifold.cpp:319-324)
    bool checkConditionsOn(const IMatrix& M, const IVector& U) {
        const bool Isolating_block = checkIsolatingBlock(M, U);
        const bool Rate_Conditions = checkRateCondition(M);
        return Isolating_block && Rate_Conditions;
    }
    // localManifold.cpp:307-325; boundDFU runs on the GPU (no CPU fallback: throws without a device)
    bool checkConditions() {
        const IVector U = constructU(U_flat_global);
        const IMatrix M = boundDFU(dimension / 2, params, F, U, stable, subdivisionNUM);
        return checkConditionsOn(M, U);
    }
    // localManifold.cpp:333-345
    IVector containmentRatios(const IVector& point_test) const {
        const int n = dimension / 2;
        IVector ratios(n);
        for (int i = 0; i < n; i++)
            ratios[i] = ia::div(point_test[i], interval(point_test[i].lo > 0 ? U_flat_global[i].hi : U_flat_global[i].lo));
        return ratios;
    }
};

// many manifolds (both ends of every front of a sweep) in ONE boundDFU launch
inline std::vector<char> checkConditionsBatch(const std::vector<localManifold*>& Ms) {
    std::vector<ccp_dfu_desc> d;
    std::vector<IVector> Us;
    for (localManifold* m : Ms) { Us.push_back(m->constructU(m->U_flat_global)); d.push_back(make_dfu_desc(m->dimension / 2, m->params, m->F, Us.back(), m->stable, m->subdivisionNUM)); }
    std::vector<char> ok(Ms.size());
    if (Ms.empty()) return ok;
    const std::vector<IMatrix> DF = boundDFUBatch(d);
    parallel_for(Ms.size(), [&](size_t q) { ok[q] = Ms[q]->checkConditionsOn(DF[q], Us[q]) ? 1 : 0; });      // per-manifold interval algebra (the pointers are distinct objects)
    return ok;
}
inline std::vector<char> checkConditionsBatch(std::vector<localManifold>& Ms) {
    std::vector<localManifold*> p;
    for (auto& m : Ms) p.push_back(&m);
    return checkConditionsBatch(p);
}

}  // namespace ccp_host
