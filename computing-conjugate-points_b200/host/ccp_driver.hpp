// ccp_driver.hpp -- host mirror of computeFrontAndConjugatePoints (heteroclinic.cpp:23-347) composed from the stage mirrors: every
// validated integration (boundary-value stage, frame propagation) and every boundDFU call runs on the GPU through the C ABI, the
// interval algebra between them on the host.  No CAPD.
//
//   (i)   manifolds: localManifold::checkConditions x2, again after resizing       heteroclinic.cpp:96-143, 203-224   ccp_manifold_conditions.hpp
//         connecting orbit: FindTime, Newton, Krawczyk, checkProof                 heteroclinic.cpp:147-281           ccp_bvp.hpp
//   (ii)  L_-: computeEigenError_minus_infty, checkConjugatePointsBelowLminus      propagateManifold.cpp:127-137      ccp_local_manifold_eig.hpp
//   (iii) frame propagation + (v) count: frameDet                                  propagateManifold.cpp:139-188      ccp_frame_det.hpp (the hot path)
//   (iv)  L_+: lastEuFrame                                                         propagateManifold.cpp:192-240      ccp_l_plus.hpp
//
// The non-rigorous preparation (approxT, getLocalGuess, coordinateChange + symplecticNormalization; heteroclinic.cpp:91-147) is the
// float shooting of ccp_setup_front.  Return values are the reference's: count >= 0, -1 manifolds / end points, -2 connecting orbit,
// -3 count failed, -4 L_+, -5 below L_-.
#pragma once
#include "ccp_l_plus.hpp"

namespace ccp_host {

struct DriverReport {
    int result = -10;
    double T = 0, radius = 0, eps_minus = 0;
    interval L_minus;
    bool conditions_first = false, conditions_resized = false, inside_box = false, proof = false, below_Lminus = false;
};

inline int computeFrontAndConjugatePoints(int dimension, const std::vector<double>& All_parameters, const IVector& All_parameters_interval,
                                          DriverReport* report = nullptr) {
    DriverReport local_rep;
    DriverReport& rep = report ? *report : local_rep;
    const int n = dimension / 2;
    // computational parameters (heteroclinic.cpp:58-87)
    const int order = 20, manifold_subdivision = 15, newton_steps = 20, grid = 14, stepsize = 5;
    const double L_minus_percent = 0.84, inflation_ratio = 1.001;
    const interval scale(dimension == 4 ? 0.000016 : 0.000001), initial_box(-.000001, .000001), L(dimension == 4 ? .00008 : .000015);

    // float preparation: eigenbasis (symplectically normalised), approximate T and boundary points in local coordinates
    ccp_front_desc d0; ccp_setup_info info;
    ccp_setup_opts o{}; o.xy_nbd = -1; o.cone_L = -1; o.eig_err = -1; o.scale = -1; o.L_minus_percent = -1; o.L_minus_override = -1;
    if (ccp_setup_front(n, All_parameters.data(), &o, &d0, &info) != 0) throw std::runtime_error(std::string("ccp: ") + ccp_last_error());
    DMatrix A(dimension, std::vector<double>(dimension));
    for (int i = 0; i < dimension; i++) for (int j = 0; j < dimension; j++) A[i][j] = d0.A_u[i * CCP_MAX_DIM + j];

    // (i) the two local manifolds (both equilibria share the eigenbasis: the Hessians agree)
    IVector U_flat(n, ia::mul(scale, interval(-1.0, 1.0)));
    localManifold localUnstable, localStable;
    localUnstable.dimension = dimension; localUnstable.stable = false; localUnstable.L = L; localUnstable.U_flat_global = U_flat;
    localUnstable.subdivisionNUM = manifold_subdivision; localUnstable.params = All_parameters_interval;
    localUnstable.F.A = A; localUnstable.F.Ainv = inverseEnclosure(toInterval(A)); localUnstable.F.p = IVector(dimension, interval(0.0));
    localStable = localUnstable; localStable.stable = true;
    for (int i = 0; i < n; i++) localStable.F.p[i] = interval(1.0);
    const IVector p_s = localStable.F.p;
    {
        std::vector<localManifold> both = {localUnstable, localStable};
        const std::vector<char> ok = checkConditionsBatch(both);
        localUnstable = both[0]; localStable = both[1];
        rep.conditions_first = ok[0] && ok[1];
        if (!rep.conditions_first) return rep.result = -1;
    }
    IVector XY_pt(dimension);
    for (int i = 0; i < n; i++) { XY_pt[i] = interval(info.x_local[i]); XY_pt[n + i] = interval(info.y_local[i]); }
    interval r_u_sqr(0.0);
    for (int i = 0; i < n; i++) r_u_sqr = ia::add(r_u_sqr, ia::sqr(XY_pt[i]));
    r_u_sqr = interval(r_u_sqr.lo);

    boundaryValueProblem BVP(n, All_parameters_interval, order, stepsize);
    boundaryValueProblem::Manifold Mu{A, localUnstable.F.p, L, false}, Ms{A, localStable.F.p, L, true};
    auto midVector = [](const IVector& v) { IVector m(v.size()); for (size_t i = 0; i < v.size(); i++) m[i] = interval(ia::mid(v[i])); return m; };
    const double T = BVP.FindTime(Mu, Ms, XY_pt, info.T);
    rep.T = T;
    const IVector XY_nbd_ZERO(dimension, interval(0.0));
    for (int i = 0; i < newton_steps; i++) XY_pt = midVector(BVP.NewtonStep(Mu, Ms, XY_pt, XY_nbd_ZERO, T, r_u_sqr));

    // resize the manifolds so that they just contain the approximate solution (heteroclinic.cpp:203-224)
    const IVector X_pt_final(XY_pt.begin(), XY_pt.begin() + n), Y_pt_final(XY_pt.begin() + n, XY_pt.end());
    const IVector x_Ratios = localUnstable.containmentRatios(X_pt_final), y_Ratios = localStable.containmentRatios(Y_pt_final);
    IVector U_flat_U_new(n), U_flat_S_new(n);
    for (int i = 0; i < n; i++) {
        U_flat_U_new[i] = ia::mul(ia::mul(U_flat[i], x_Ratios[i]), interval(inflation_ratio));
        U_flat_S_new[i] = ia::mul(ia::mul(U_flat[i], y_Ratios[i]), interval(inflation_ratio));
    }
    localUnstable.U_flat_global = U_flat_U_new; localStable.U_flat_global = U_flat_S_new;          // updateSize
    {
        std::vector<localManifold> both = {localUnstable, localStable};
        const std::vector<char> ok = checkConditionsBatch(both);
        localUnstable = both[0]; localStable = both[1];
        rep.conditions_resized = ok[0] && ok[1];
        if (!rep.conditions_resized) return rep.result = -1;
    }

    // Krawczyk validation of the connecting orbit (heteroclinic.cpp:229-281)
    IVector XY_nbd(dimension, ia::mul(scale, initial_box));
    for (int i = 0; i < newton_steps; i++) {
        const IVector out = BVP.NewtonStep(Mu, Ms, XY_pt, XY_nbd, T, r_u_sqr);
        XY_pt = midVector(out);
        for (int k = 0; k < dimension; k++) XY_nbd[k] = ia::sub(out[k], XY_pt[k]);
    }
    for (int k = 0; k < dimension; k++) XY_nbd[k] = ia::mul(XY_nbd[k], interval(1.5));
    BVP.NewtonStep(Mu, Ms, XY_pt, XY_nbd, T, r_u_sqr);
    rep.inside_box = true;
    for (int i = 0; i < n; i++) {
        rep.inside_box = rep.inside_box && ia::subsetInterior(ia::add(XY_pt[i], XY_nbd[i]), U_flat_U_new[i])
                                        && ia::subsetInterior(ia::add(XY_pt[n + i], XY_nbd[n + i]), U_flat_S_new[i]);
        rep.radius = std::max(rep.radius, std::max(ia::mag(XY_nbd[i]), ia::mag(XY_nbd[n + i])));
    }
    if (!rep.inside_box) return rep.result = -1;
    rep.proof = BVP.checkProof();
    if (!rep.proof) return rep.result = -2;

    // (ii) L_- and the eigenfunction error at minus infinity
    localManifold_Eig localUnstable_eig(localUnstable);
    localUnstable_eig.computeEigenError_minus_infty();
    rep.eps_minus = localUnstable_eig.eps_unscaled.hi;
    rep.below_Lminus = localUnstable_eig.checkConjugatePointsBelowLminus();

    // (iii)-(v) the hot path with the L_- and L_+ stages plugged into the frameDet shim
    const IVector X_pt(XY_pt.begin(), XY_pt.begin() + n), X_nbd(XY_nbd.begin(), XY_nbd.begin() + n);
    const IVector Y_pt(XY_pt.begin() + n, XY_pt.end()), Y_nbd(XY_nbd.begin() + n, XY_nbd.end());
    const FlowSet start = boundaryValueProblem::getPointNbd(Mu, X_pt, X_nbd);
    propagateManifold E_u(n, All_parameters_interval, A, start.p_i, start.Uxy, localUnstable_eig.Eu_m_Error_Final, order, stepsize);
    const interval L_minus(ia::mul(ia::mul(interval(2.0), interval(T)), interval(L_minus_percent)).hi);
    rep.L_minus = L_minus;
    const double effective_L_plus = 2 * T * (1 - L_minus_percent);
    IVector endPoint_LPlus = BVP.Gxy(boundaryValueProblem::getPointNbd(Ms, Y_pt, Y_nbd), effective_L_plus);
    for (int i = 0; i < dimension; i++) endPoint_LPlus[i] = ia::sub(endPoint_LPlus[i], p_s[i]);
    LPlusStage st;
    st.dimension = dimension; st.manifold_subdivision = manifold_subdivision; st.stable = localStable;
    E_u.below_Lminus = [&] { return rep.below_Lminus; };
    E_u.L_plus = [&](const topFrame& f, const IVector& e) { return st.lastEuFrame(f, e); };
    return rep.result = E_u.frameDet(L_minus, grid, endPoint_LPlus);
}

}  // namespace ccp_host
