// ccp_l_plus.hpp -- host mirror of the L_+ stage that consumes the hot path's output (propagateManifold.cpp:192-558).
//
//   bool     propagateManifold::lastEuFrame(topFrame&, IVector endPoint_LPlus)             :192-240   reads A_frame.getLastFrame()
//   localManifold_Eig construct_Manifold_at_LPlus(IVector endPoint_LPlus)                  :244-304   up to 10 checkConditions (boundDFU on the GPU)
//   IMatrix  projectionGammaBeta(IMatrix& last_Frame, const IMatrix& EFunction_Error)      :308-372
//   bool     checkL_plus(topFrame&, IMatrix U_coord, interval eps_0, IVector eigenvalues)  :375-432
//   bool     checkL_plus_local(Gamma, Beta, eps_0, nu_1, nu_n)                             :434-510   Proposition 2.9
//   interval compute_epsilon_beta(Gamma, Beta)                                             :512-558   Section 3.4
//   bool     topFrame::checkFirstFrame(int j_hat)                                          topFrame.cpp:336-351
//
// Plug `LPlusStage::lastEuFrame` into `propagateManifold::L_plus` (ccp_frame_det.hpp) and frameDet returns the reference's -4 / count
// without CAPD.  gauss(M, b) is inverseEnclosure(M) * b, CAPD's det is a cofactor expansion, norms as in ccp_manifold_conditions.hpp.
#pragma once
#include "ccp_local_manifold_eig.hpp"

namespace ccp_host {

inline interval detCofactor(const IMatrix& M) {
    const int n = (int)M.size();
    if (n == 1) return M[0][0];
    if (n == 2) return ia::sub(ia::mul(M[0][0], M[1][1]), ia::mul(M[0][1], M[1][0]));
    interval d(0.0);
    for (int c = 0; c < n; c++) {
        IMatrix m(n - 1, IVector(n - 1));
        for (int i = 1; i < n; i++) { int cc = 0; for (int j = 0; j < n; j++) if (j != c) m[i - 1][cc++] = M[i][j]; }
        const interval t = ia::mul(M[0][c], detCofactor(m));
        d = (c % 2 == 0) ? ia::add(d, t) : ia::sub(d, t);
    }
    return d;
}
inline IMatrix transposeM(const IMatrix& A) {
    IMatrix T(A[0].size(), IVector(A.size()));
    for (size_t i = 0; i < A.size(); i++) for (size_t j = 0; j < A[0].size(); j++) T[j][i] = A[i][j];
    return T;
}
inline IVector matVec(const IMatrix& A, const IVector& x) {
    IVector y(A.size());
    for (size_t i = 0; i < A.size(); i++) { interval a(0.0); for (size_t j = 0; j < x.size(); j++) a = ia::add(a, ia::mul(A[i][j], x[j])); y[i] = a; }
    return y;
}
inline interval intervalHull(const interval& a, const interval& b) { return interval(std::min(a.lo, b.lo), std::max(a.hi, b.hi)); }

// topFrame.cpp:336-351: F = getFirstFrame(j_hat) has full rank if det(F^T F) is bounded away from zero
inline bool checkFirstFrame(const topFrame& A_frame, int j_hat) {
    const IMatrix F = A_frame.getFirstFrame(j_hat);
    const interval det_val = detCofactor(matMul(transposeM(F), F));
    return ia::abs(det_val).lo > 0;
}

// propagateManifold.cpp:512-558
inline interval compute_epsilon_beta(const IMatrix& Gamma, const IMatrix& Beta) {
    IMatrix Gc = Gamma, Gd = Gamma;
    for (size_t i = 0; i < Gamma.size(); i++) for (size_t j = 0; j < Gamma[0].size(); j++) {
        Gc[i][j] = interval(ia::mid(Gamma[i][j])); Gd[i][j] = ia::sub(Gamma[i][j], Gc[i][j]);
    }
    const IMatrix GcT = transposeM(Gc), GdT = transposeM(Gd);
    const interval mu_c = ml(matMul(GcT, Gc));
    const interval mu = ia::sub(ia::sub(mu_c, ia::mul(interval(2.0), euclNorm(matMul(GdT, Gc)))), euclNorm(matMul(GdT, Gd)));
    if (mu.lo < 0) return interval(-1.0);
    return ia::div(euclNorm(Beta), ia::sqrt(mu));
}

// propagateManifold.cpp:434-510 (diffusion matrix D = identity: d_min = d_max = 1)
inline bool checkL_plus_local(const IMatrix& Gamma, const IMatrix& Beta, interval eps_0, const interval& nu_1, const interval& nu_n, int dimension) {
    interval epsilon_beta = compute_epsilon_beta(Gamma, Beta);
    if (epsilon_beta.hi < 0) return false;
    epsilon_beta = interval(epsilon_beta.hi);
    eps_0 = interval(eps_0.hi);
    const interval one(1.0), two(2.0), n((double)(dimension / 2)), d_min(1.0), d_max(1.0);
    const interval s2n1 = ia::sqrt(ia::mul(ia::mul(ia::mul(two, n), nu_1), d_max));                 // sqrt(2 n nu_1 d_max)
    const interval sum_for_neumann_test = ia::mul(eps_0, s2n1);
    if (!(sum_for_neumann_test.hi < 1)) return false;                                                // the reference tests this after C_P; here first: 1 - eps_0 s must not reach 0
    interval C_P = ia::div(ia::mul(two, s2n1), ia::sub(one, sum_for_neumann_test));
    interval C_Q = ia::add(ia::add(ia::sqrt(ia::div(ia::mul(two, n), ia::mul(nu_n, d_min))), s2n1), ia::mul(ia::mul(two, eps_0), n));
    C_P = interval(C_P.hi); C_Q = interval(C_Q.hi);
    const interval eQ = ia::mul(eps_0, C_Q), eP = ia::mul(eps_0, C_P);
    const interval C_M1 = ia::mul(eQ, ia::add(two, eQ));
    const interval C_M2 = ia::add(ia::add(ia::mul(eP, ia::add(two, eP)), ia::mul(ia::mul(two, epsilon_beta), ia::add(one, eP))), ia::sqr(epsilon_beta));
    const bool NEUMANN_SERIES_TEST = C_M1.hi < 1 && C_M2.hi < 1;
    if (!NEUMANN_SERIES_TEST) return false;                                                          // (also keeps 1 - C_M away from zero below)
    const interval sum_1 = ia::mul(ia::mul(two, eps_0), ia::add(C_Q, C_P));
    const interval sum_2 = ia::mul(ia::sqr(eps_0), ia::add(ia::sqr(C_Q), ia::sqr(C_P)));
    const interval sum_3 = ia::add(ia::sqr(epsilon_beta), ia::mul(ia::mul(two, epsilon_beta), ia::add(one, eP)));
    const interval sum_4 = ia::div(ia::mul(ia::sqr(ia::add(one, eQ)), C_M1), ia::abs(ia::sub(one, C_M1)));
    const interval sum_5 = ia::div(ia::mul(ia::sqr(ia::add(ia::add(one, eP), epsilon_beta)), C_M2), ia::abs(ia::sub(one, C_M2)));
    const interval total_sum = ia::add(ia::add(ia::add(ia::add(sum_1, sum_2), sum_3), sum_4), sum_5);
    return total_sum.hi < 1;
}

struct LPlusStage {
    int dimension = 6;
    int manifold_subdivision = 15;
    localManifold stable;                 // *pStable: its F (A_s, Ainv, p) and the parameters

    // propagateManifold.cpp:244-304
    // the manifold object before the adjust_L loop (:244-289)
    localManifold_Eig start_Manifold_at_LPlus(IVector endPoint_LPlus) const {
        const int n = dimension / 2;
        endPoint_LPlus = matVec(stable.F.Ainv, endPoint_LPlus);                                     // gauss(A_s, .): into eigen-coordinates
        IVector vec_Lplus_u(endPoint_LPlus.begin(), endPoint_LPlus.begin() + n), vec_Lplus_s(endPoint_LPlus.begin() + n, endPoint_LPlus.end());
        IVector U_flat_new(n);
        for (int i = 0; i < n; i++) U_flat_new[i] = ia::mul(vec_Lplus_s[i], interval(-1e-10, 1 + 1e-10));
        interval ss(0.0);
        for (const interval& x : vec_Lplus_s) ss = ia::add(ss, ia::sqr(x));
        const interval L_new(ia::div(euclNorm(vec_Lplus_u), interval(ia::sqrt(ss).lo)).hi);           // right end of |u| / |s|  (vartheta of the paper)
        localManifold_Eig big(stable);
        big.stable = true; big.U_flat_global = U_flat_new; big.L = L_new; big.subdivisionNUM = manifold_subdivision;
        return big;
    }
    localManifold_Eig construct_Manifold_at_LPlus(const IVector& endPoint_LPlus) const {
        localManifold_Eig big = start_Manifold_at_LPlus(endPoint_LPlus);
        for (int i = 0; i < 10; i++) {                                                              // adjust_L
            if (big.checkConditions()) break;
            big.L = ia::mul(big.L, interval(2.0));
        }
        return big;
    }

    // propagateManifold.cpp:308-372
    IMatrix projectionGammaBeta(const IMatrix& last_Frame, const IMatrix& EFunction_Error) const {
        const int n = dimension / 2;
        IMatrix M = matMul(stable.F.Ainv, EFunction_Error);                                         // eye + krawczykInverse(A_s) * EFunction_Error
        for (int i = 0; i < dimension; i++) M[i][i] = ia::add(M[i][i], interval(1.0));
        const IMatrix Minv = inverseEnclosure(M);
        IMatrix U_coord(dimension, IVector(n));
        for (int j = 0; j < n; j++) {
            IVector col(dimension);
            for (int i = 0; i < dimension; i++) col[i] = last_Frame[i][j];
            IVector v = matVec(Minv, matVec(stable.F.Ainv, col));
            IVector av(dimension);
            for (int i = 0; i < dimension; i++) av[i] = ia::abs(v[i]);
            const interval mx = getMax(av);
            for (int i = 0; i < dimension; i++) U_coord[i][j] = ia::div(v[i], mx);
        }
        return U_coord;
    }

    // propagateManifold.cpp:375-432
    bool checkL_plus(const topFrame& A_frame, const IMatrix& U_coord, const interval& eps_0, const IVector& eigenvalues) const {
        const int n = dimension / 2;
        const interval nu_1 = eigenvalues[0], nu_n = eigenvalues[n - 1];
        for (int k = 0; k < n; k++) {                                                               // remove k, the hat-index
            IMatrix Gamma(n, IVector(n - 1)), Beta(n, IVector(n - 1));
            int c = 0;
            for (int j = 0; j < n; j++) {
                if (j == k) continue;
                for (int i = 0; i < n; i++) { Gamma[i][c] = U_coord[i][j]; Beta[i][c] = U_coord[i + n][j]; }
                c++;
            }
            if (!checkFirstFrame(A_frame, k)) continue;
            if (checkL_plus_local(Gamma, Beta, eps_0, nu_1, nu_n, dimension)) return true;
        }
        return false;
    }

    // propagateManifold.cpp:192-240
    bool lastEuFrame(const topFrame& A_frame, const IVector& endPoint_LPlus) const {
        localManifold_Eig big = construct_Manifold_at_LPlus(endPoint_LPlus);
        return lastEuFrameWith(A_frame, big, big.checkConditions());
    }
    // the part after the manifold at L_+ has been built and checked (:206-239)
    bool lastEuFrameWith(const topFrame& A_frame, localManifold_Eig& big, bool conditions_S) const {
        const IMatrix last_Frame = A_frame.getLastFrame();
        if (!conditions_S) return false;
        big.computeEigenError_plus_infty();
        const IMatrix& EFunction_Error = big.Eigenfunction_Error_plus_infty;
        interval eps_0(0.0);
        for (int i = 0; i < dimension; i++) eps_0 = intervalHull(eps_0, EFunction_Error[0][i]);
        const IMatrix U_coord = projectionGammaBeta(last_Frame, EFunction_Error);
        return checkL_plus(A_frame, U_coord, eps_0, big.eigenvalues);
    }
};

// construct_Manifold_at_LPlus for many fronts: every round of the adjust_L loop is ONE boundDFU launch over the manifolds not yet verified.
// Returns the manifolds and whether each was verified (what the final checkConditions of lastEuFrame would return).
inline std::vector<localManifold_Eig> constructManifoldsAtLPlusBatch(const std::vector<LPlusStage>& st, const std::vector<IVector>& endPoints,
                                                                      std::vector<char>& verified) {
    std::vector<localManifold_Eig> big;
    for (size_t q = 0; q < st.size(); q++) big.push_back(st[q].start_Manifold_at_LPlus(endPoints[q]));
    verified.assign(st.size(), 0);
    for (int i = 0; i < 11; i++) {                                                                  // 10 adjust_L rounds + the final check
        std::vector<localManifold*> todo; std::vector<size_t> who;
        for (size_t q = 0; q < big.size(); q++) if (!verified[q]) { todo.push_back(&big[q]); who.push_back(q); }
        if (todo.empty()) break;
        const std::vector<char> ok = checkConditionsBatch(todo);
        for (size_t k = 0; k < who.size(); k++) {
            if (ok[k]) verified[who[k]] = 1;
            else if (i < 10) big[who[k]].L = ia::mul(big[who[k]].L, interval(2.0));                 // round 10 is lastEuFrame's own check: no doubling
        }
    }
    return big;
}

}  // namespace ccp_host
