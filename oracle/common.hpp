// oracle/common.hpp -- TEST INFRASTRUCTURE (CPU oracle).  Not part of the product path.
//
// Pieces shared by both oracle modes: the n x n interval determinant, Jacobi's formula and the
// per-sub-interval decision tree of topFrame::countZeros (reference topFrame.cpp:193-263),
// the sub-interval grid of getTotalTrajectory (utils.cpp:219-232,242).
#pragma once
#include "ival.hpp"
#include "../include/ccp.h"
#include <vector>
#include <cstring>

namespace orc {

typedef ival Frame[CCP_MAX_N][CCP_MAX_N];

// det(IMatrix) -- CAPD uses interval Gaussian elimination [CAPD-recollection]; cofactor expansion
// along the first row is used here (rigorous, no division; SURVEY.md 7 "Determinant enclosure").
static inline ival det2(const ival& a, const ival& b, const ival& c, const ival& d) { return a * d - b * c; }
static inline ival det3(const ival m[3][3]) {
    return m[0][0] * det2(m[1][1], m[1][2], m[2][1], m[2][2])
         - m[0][1] * det2(m[1][0], m[1][2], m[2][0], m[2][2])
         + m[0][2] * det2(m[1][0], m[1][1], m[2][0], m[2][1]);
}
static inline ival detN(const Frame F, int n);
static inline ival minor_det(const Frame F, int n, int ih, int jh) {   // topFrame::minorDet (topFrame.cpp:124-153)
    Frame Mn;
    for (int i = 0; i < n - 1; i++) for (int j = 0; j < n - 1; j++) Mn[i][j] = F[i < ih ? i : i + 1][j < jh ? j : j + 1];
    return detN(Mn, n - 1);
}
static inline ival detN(const Frame F, int n) {
    if (n == 1) return F[0][0];
    if (n == 2) return det2(F[0][0], F[0][1], F[1][0], F[1][1]);
    if (n == 3) { ival m[3][3]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = F[i][j]; return det3(m); }
    ival acc = F[0][0] * minor_det(F, n, 0, 0);
    for (int j = 1; j < n; j++) {
        ival term = F[0][j] * minor_det(F, n, 0, j);
        acc = (j & 1) ? acc - term : acc + term;
    }
    return acc;
}
// topFrame::calculateDerivative (topFrame.cpp:110-122): trace(adj(A)*A'), adj = transpose(sign*minor)
static inline ival jacobi_derivative(const Frame A, const Frame Ad, int n) {
    ival tr(0.0);
    for (int i = 0; i < n; i++) {
        ival diag(0.0);
        for (int j = 0; j < n; j++) {
            ival cof = minor_det(A, n, j, i);          // adj[i][j] = sign(j,i)*minor(j,i)
            if ((i + j) & 1) cof = -cof;
            diag = diag + cof * Ad[j][i];
        }
        tr = tr + diag;
    }
    return tr;
}

struct Counter {
    int zero_count = 0, failure_count = 0, first_failure = -1;
    int n_conj = 0;
    ccp_interval conj_time[CCP_MAX_CONJ];
    int conj_index[CCP_MAX_CONJ];
};

// One pass of the loop body of topFrame::countZeros (topFrame.cpp:206-254).
// L,R,V = det of left/right/range frame, D = Jacobi derivative over the range, time = prevTime+sub.
static inline void count_step(Counter& c, int index, const ival& time, const ival& L, const ival& R, const ival& V, const ival& D) {
    ival LR = L * R;
    bool fail = false;
    if (LR.lo > 0) {
        if (definite(V)) return;
        if (definite(D)) return;
        ival w = time - ival(time.lo);                       // time_series[i]-time_series[i].left()
        ival mvt = w * D;
        ival half(mvt.lo * 0.5, mvt.hi * 0.5);               // MVT_deriv/2 (exact)
        ival bound = hull(L + half, R - half);
        if (definite(bound)) return;
        fail = true;
    } else if (LR.hi < 0) {
        if (definite(D)) {
            if (c.n_conj < CCP_MAX_CONJ) { c.conj_time[c.n_conj].lo = time.lo; c.conj_time[c.n_conj].hi = time.hi; c.conj_index[c.n_conj] = index; c.n_conj++; }
            c.zero_count++;
            return;
        }
        fail = true;
    } else fail = true;
    if (fail) { if (c.failure_count == 0) c.first_failure = index; c.failure_count++; }
}

// sub-interval end points of one step (utils.cpp:219-223): gp[i] = inf(i*h/grid), gp[grid] = h
static inline void grid_points(double h, int grid, double* gp) {
    for (int i = 0; i < grid; i++) gp[i] = div_rd(mul_rd((double)i, h), (double)grid);
    gp[grid] = h;
}

static inline int step_count(double L_minus, double h) {
    // fixed step, last one shortened to land on L_minus exactly (ITimeMap with stopAfterStep, utils.cpp:204-258)
    int S = (int)std::ceil(L_minus / h);
    while ((double)(S - 1) * h >= L_minus) S--;
    while ((double)S * h < L_minus) S++;
    return S;
}

} // namespace orc
