// ccp_local_manifold.hpp -- C++ host shim for localManifold::boundDFU (localManifold.cpp:347-409) above ccp_bound_dfu_batch.
//
//   IMatrix localManifold::boundDFU(IVector U)        hull over the N^n sub-boxes of U of  (*pF)[U_part] = Ainv * Df(p + A U_part) * A
//
// The members it reads from the CAPD objects:  (*pF).A, (*pF).Ainv, (*pF).p (localVField.h:17-19), `stable`, `subdivisionNUM`
// (localManifold.h:28-36) and the parameters.  checkConditions (localManifold.cpp:307-325) calls it once per check, the driver at
// least four times per front (heteroclinic.cpp:113-224) and construct_Manifold_at_LPlus up to ten more (propagateManifold.cpp:293-303);
// `boundDFUBatch` takes all of them (or a whole sweep) in one launch.
#pragma once
#include "ccp_frame_det.hpp"

namespace ccp_host {

struct LocalField {              // what localVField holds
    DMatrix A;                   // point matrix (mid-matrix, heteroclinic.cpp:116-117)
    IMatrix Ainv;                // gaussInverseMatrix(A), stays CAPD
    IVector p;                   // equilibrium
};

inline ccp_dfu_desc make_dfu_desc(int n, const IVector& params, const LocalField& F, const IVector& U, bool stable, int subdivisionNUM) {
    if (n < 2 || n > CCP_MAX_N) throw std::invalid_argument("n out of range");
    ccp_dfu_desc d{};
    d.n = n; d.subdivision = subdivisionNUM; d.stable = stable ? 1 : 0;
    for (int i = 0; i < n; i++) d.b[i] = to_c(params[i]);
    for (int i = 0; i < n - 1; i++) d.c[i] = to_c(params[n + i]);
    for (int i = 0; i < 2 * n; i++) {
        d.p[i] = to_c(F.p[i]); d.U[i] = to_c(U[i]);
        for (int j = 0; j < 2 * n; j++) { d.A[i * CCP_MAX_DIM + j] = F.A[i][j]; d.Ainv[i * CCP_MAX_DIM + j] = to_c(F.Ainv[i][j]); }
    }
    return d;
}

inline std::vector<IMatrix> boundDFUBatch(const std::vector<ccp_dfu_desc>& descs) {
    std::vector<ccp_dfu_result> r(descs.size());
    if (ccp_bound_dfu_batch(descs.data(), descs.size(), r.data()) != 0) throw std::runtime_error(std::string("ccp: ") + ccp_last_error());
    std::vector<IMatrix> out(descs.size());
    for (size_t q = 0; q < descs.size(); q++) {
        if (r[q].status != CCP_OK) throw std::runtime_error("ccp: boundDFU status " + std::to_string(r[q].status));
        const int D = 2 * descs[q].n;
        out[q].assign(D, IVector(D));
        for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) out[q][i][j] = from_c(r[q].DF[i * CCP_MAX_DIM + j]);
    }
    return out;
}

// drop-in for `IMatrix M = boundDFU(U);` (localManifold.cpp:316)
inline IMatrix boundDFU(int n, const IVector& params, const LocalField& F, const IVector& U, bool stable, int subdivisionNUM) {
    return boundDFUBatch({make_dfu_desc(n, params, F, U, stable, subdivisionNUM)})[0];
}

} // namespace ccp_host
