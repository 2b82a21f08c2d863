// ccp_bvp.hpp -- C++ host shim for the validated integrations of the boundary-value stage, above the same C ABI call.
//
//   IVector boundaryValueProblem::Gxy (XY_pt, XY_nbd, T, STABLE)     boundaryValueProblem.cpp:9-59
//       = Phi(T, C0Rect2Set(p_i, A_i, Uxy))            (p_i, Uxy) = pManifold->getPointNbd(XY_pt, XY_nbd), A_i = (*pManifold->pF).A
//   IMatrix boundaryValueProblem::DGxy(XY_pt, XY_nbd, T, STABLE)     boundaryValueProblem.cpp:61-115
//       = XY_deriv * A_i * pManifold->DW               XY_deriv from Phi(T, C1Rect2Set(p_i, A_i, Uxy), XY_deriv)
//
// Both are the C^1 time-T map of the front ODE u' = v, v' = -grad G(u) over the set p_i + A_i*Uxy, which the engine computes anyway
// while it propagates the frame: the image is column n+1 of ccp_front_result::last_frame, [D Phi_T] A_i is ccp_front_result::flow_P.
// 41 Newton / Krawczyk steps x 4 such integrations per front (heteroclinic.cpp:171-251) go through here; the Newton logic itself
// (NewtonStep, krawczykInverse, FindTime) stays host C++ in the reference and multiplies flow_P by DW (boundaryValueProblem.cpp:111).
//
// STABLE = true integrates the reversed field f_minus from the stable manifold (boundaryValueProblem.cpp:22-32).  The front ODE is
// reversible: with S = diag(I, -I),  Phi^minus_T = S Phi_T S,  so the backward maps are the forward maps of the mirrored set
// (S p_i, S A_i, Uxy), mirrored back:  image -> S image,  [D Phi^minus_T] A_i = S [D Phi_T (S A_i)].
//
// Integration parameters are the engine's (fixed step 2^-step_size, order); T is a point (the reference passes a thin interval).
#pragma once
#include "ccp_frame_det.hpp"

namespace ccp_host {

struct FlowSet {                 // what getPointNbd / (*pF).A hand to C0Rect2Set / C1Rect2Set
    IVector p_i, Uxy;            // 2n each
    DMatrix A_i;                 // 2n x 2n point matrix
    bool stable = false;         // true: backward flow from the stable manifold (f_minus)
};

class boundaryValueProblem {
  public:
    boundaryValueProblem(int n, const IVector& params, int order, int step_size) : n_(n), params_(params), order_(order), step_(step_size) {
        if (n < 2 || n > CCP_MAX_N) throw std::invalid_argument("n out of range");
    }
    // enclosure of the time-T image of the set (Gxy)
    IVector Gxy(const FlowSet& s, double T) { return run({s}, {T})[0].image; }
    // enclosure of [D Phi_T] * A_i over the set (DGxy before the factor DW)
    IMatrix DGxyA(const FlowSet& s, double T) { return run({s}, {T})[0].deriv_A; }

    struct Out { int status; IVector image; IMatrix deriv_A; };
    // all integrations of a Newton step (or of a whole sweep) in ONE launch
    std::vector<Out> run(const std::vector<FlowSet>& sets, const std::vector<double>& T) {
        const int D = 2 * n_;
        std::vector<ccp_front_desc> d(sets.size());
        std::vector<ccp_front_result> r(sets.size());
        for (size_t q = 0; q < sets.size(); q++) {
            const FlowSet& s = sets[q];
            ccp_front_desc& e = d[q];
            e = ccp_front_desc{};
            e.n = n_; e.order = order_; e.step_log2 = step_; e.grid = 1; e.L_minus = T[q];
            for (int i = 0; i < n_; i++) e.b[i] = to_c(params_[i]);
            for (int i = 0; i < n_ - 1; i++) e.c[i] = to_c(params_[n_ + i]);
            for (int i = 0; i < D; i++) {
                const double sg = (s.stable && i >= n_) ? -1.0 : 1.0;                       // S = diag(I, -I)
                e.p_i[i] = sg > 0 ? to_c(s.p_i[i]) : to_c(interval(-s.p_i[i].hi, -s.p_i[i].lo));
                e.Uxy[i] = to_c(s.Uxy[i]);
                for (int j = 0; j < D; j++) e.A_u[i * CCP_MAX_DIM + j] = sg * s.A_i[i][j];
            }
        }
        if (ccp_frame_count_batch(d.data(), d.size(), r.data()) != 0) throw std::runtime_error(std::string("ccp: ") + ccp_last_error());
        std::vector<Out> out(sets.size());
        for (size_t q = 0; q < sets.size(); q++) {
            Out& o = out[q];
            o.status = r[q].status;
            o.image.resize(D); o.deriv_A.assign(D, IVector(D));
            for (int i = 0; i < D; i++) {
                const bool flip = sets[q].stable && i >= n_;
                auto m = [&](const ccp_interval& a) { return flip ? interval(-a.hi, -a.lo) : from_c(a); };
                o.image[i] = m(r[q].last_frame[i * (CCP_MAX_N + 2) + n_ + 1]);
                for (int j = 0; j < D; j++) o.deriv_A[i][j] = m(r[q].flow_P[i * CCP_MAX_DIM + j]);
            }
        }
        return out;
    }

  private:
    int n_; IVector params_; int order_, step_;
};

} // namespace ccp_host
