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
#include <chrono>
#include "ccp_l_plus.hpp"
#include "../../include/ccp_setup.h"

namespace ccp_host {

struct DriverReport {
    int result = -10;
    double T = 0, radius = 0, eps_minus = 0;
    int zero_count = -1, failure_count = -1, first_failure = -1, series_length = 0;      // countZeros of the frame stage
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

    boundaryValueProblem BVP(n, All_parameters_interval, order);                  // adaptive in the reference: step ladder from 2^-3 (ccp_bvp.hpp)
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
    rep.result = E_u.frameDet(L_minus, grid, endPoint_LPlus);
    const ccp_front_result& fr = E_u.frame().result();
    rep.zero_count = fr.zero_count; rep.failure_count = fr.failure_count; rep.first_failure = fr.first_failure; rep.series_length = fr.series_length;
    return rep.result;
}

// ---- the same for a whole parameter sweep (SampleParameters, heteroclinic.cpp:429-491): all fronts advance through the stages in
// lock-step, so every stage is ONE launch over the fronts still alive (about 65 launches for any number of fronts) instead of
// about 65 launches per front.  Per front the operations and results are those of computeFrontAndConjugatePoints above; a front that
// fails a stage keeps the reference's code for that stage, an exception inside its host algebra gives -10 (or the thrown int) as in
// SampleParameters' try / catch (heteroclinic.cpp:456-467).
// wall-clock seconds of the stages of the last computeFrontAndConjugatePointsBatch call (stage name, seconds), for the timing reports
inline std::vector<std::pair<std::string, double>>& driverStageTimes() { static std::vector<std::pair<std::string, double>> t; return t; }
inline std::vector<int> computeFrontAndConjugatePointsBatch(int dimension, const std::vector<std::vector<double>>& P, const std::vector<IVector>& PI,
                                                           std::vector<DriverReport>* reports = nullptr) {
    auto& stage_times = driverStageTimes(); stage_times.clear();
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* name) { const auto now = std::chrono::steady_clock::now(); stage_times.emplace_back(name, std::chrono::duration<double>(now - t_last).count()); t_last = now; };
    const int n = dimension / 2;
    const size_t m = P.size();
    const int order = 20, manifold_subdivision = 15, newton_steps = 20, grid = 14, stepsize = 5;
    const double L_minus_percent = 0.84, inflation_ratio = 1.001;
    const interval scale(dimension == 4 ? 0.000016 : 0.000001), initial_box(-.000001, .000001), L(dimension == 4 ? .00008 : .000015);
    const int ALIVE = -1000;
    std::vector<int> code(m, ALIVE);
    std::vector<DriverReport> rep(m);
    struct Front {
        DMatrix A; localManifold U, S; boundaryValueProblem::Manifold Mu, Ms; IVector XY_pt, XY_nbd, U_new, S_new, endPoint; interval r_u_sqr, L_minus;
        double T = 0; bool success = false; localManifold_Eig eig; FlowSet start;
    };
    std::vector<Front> F(m);
    auto alive = [&](size_t q) { return code[q] == ALIVE; };
    auto guarded = [&](size_t q, const std::function<void()>& fn) {            // SampleParameters' try / catch around one front
        try { fn(); } catch (const std::exception&) { code[q] = -10; } catch (int e) { code[q] = e; }
    };
    auto midVector = [](const IVector& v) { IVector o(v.size()); for (size_t i = 0; i < v.size(); i++) o[i] = interval(ia::mid(v[i])); return o; };

    // float preparation for all fronts
    std::vector<double> flat;
    for (const auto& p : P) flat.insert(flat.end(), p.begin(), p.end());
    std::vector<ccp_front_desc> d0(m); std::vector<ccp_setup_info> info(m);
    ccp_setup_opts o{}; o.xy_nbd = -1; o.cone_L = -1; o.eig_err = -1; o.scale = -1; o.L_minus_percent = -1; o.L_minus_override = -1;
    if (ccp_setup_sweep(n, flat.data(), m, &o, d0.data(), info.data(), 0) != 0) throw std::runtime_error(std::string("ccp: ") + ccp_last_error());

    lap("float preparation (ccp_setup_sweep)");
    const IVector U_flat(n, ia::mul(scale, interval(-1.0, 1.0)));
    for (size_t q = 0; q < m; q++) guarded(q, [&] {
        Front& f = F[q];
        f.A.assign(dimension, std::vector<double>(dimension));
        for (int i = 0; i < dimension; i++) for (int j = 0; j < dimension; j++) f.A[i][j] = d0[q].A_u[i * CCP_MAX_DIM + j];
        f.U.dimension = dimension; f.U.stable = false; f.U.L = L; f.U.U_flat_global = U_flat; f.U.subdivisionNUM = manifold_subdivision; f.U.params = PI[q];
        f.U.F.A = f.A; f.U.F.Ainv = inverseEnclosure(toInterval(f.A)); f.U.F.p = IVector(dimension, interval(0.0));
        f.S = f.U; f.S.stable = true;
        for (int i = 0; i < n; i++) f.S.F.p[i] = interval(1.0);
        f.Mu = boundaryValueProblem::Manifold{f.A, f.U.F.p, L, false}; f.Ms = boundaryValueProblem::Manifold{f.A, f.S.F.p, L, true};
        f.XY_pt.resize(dimension);
        for (int i = 0; i < n; i++) { f.XY_pt[i] = interval(info[q].x_local[i]); f.XY_pt[n + i] = interval(info[q].y_local[i]); }
        interval r(0.0);
        for (int i = 0; i < n; i++) r = ia::add(r, ia::sqr(f.XY_pt[i]));
        f.r_u_sqr = interval(r.lo);
        f.T = info[q].T;
    });
    auto checkBoth = [&](bool DriverReport::*flag) {
        std::vector<localManifold*> ms; std::vector<size_t> who;
        for (size_t q = 0; q < m; q++) if (alive(q)) { ms.push_back(&F[q].U); ms.push_back(&F[q].S); who.push_back(q); }
        const std::vector<char> ok = checkConditionsBatch(ms);
        for (size_t k = 0; k < who.size(); k++) { rep[who[k]].*flag = ok[2 * k] && ok[2 * k + 1]; if (!(rep[who[k]].*flag)) code[who[k]] = -1; }
    };
    lap("front set-up");
    checkBoth(&DriverReport::conditions_first);
    lap("manifold conditions, first");

    boundaryValueProblem BVP(n, PI.empty() ? IVector(2 * n - 1, interval(0.0)) : PI[0], order);
    {   // FindTime
        std::vector<boundaryValueProblem::TimeIn> in; std::vector<size_t> who;
        for (size_t q = 0; q < m; q++) if (alive(q)) { in.push_back({F[q].Mu, F[q].Ms, F[q].XY_pt, F[q].T, PI[q]}); who.push_back(q); }
        const std::vector<double> T = BVP.FindTimeBatch(in);
        for (size_t k = 0; k < who.size(); k++) { F[who[k]].T = T[k]; rep[who[k]].T = T[k]; }
    }
    auto newtonRound = [&](bool zero_nbd) {
        std::vector<boundaryValueProblem::StepIn> in; std::vector<size_t> who;
        for (size_t q = 0; q < m; q++) if (alive(q)) {
            in.push_back({F[q].Mu, F[q].Ms, F[q].XY_pt, zero_nbd ? IVector(dimension, interval(0.0)) : F[q].XY_nbd, F[q].T, F[q].r_u_sqr, PI[q]});
            who.push_back(q);
        }
        const std::vector<boundaryValueProblem::StepOut> out = BVP.NewtonStepBatch(in);
        for (size_t k = 0; k < who.size(); k++) {
            Front& f = F[who[k]];
            f.success = out[k].success;
            if (zero_nbd) f.XY_pt = midVector(out[k].kraw_image);
            else { const IVector mid = midVector(out[k].kraw_image); for (int i = 0; i < dimension; i++) f.XY_nbd[i] = ia::sub(out[k].kraw_image[i], mid[i]); f.XY_pt = mid; }
        }
    };
    lap("FindTime");
    for (int i = 0; i < newton_steps; i++) newtonRound(true);
    lap("Newton steps");
    stage_times.emplace_back("  (validated-integration launches so far)", (double)BVP.launches());
    stage_times.emplace_back("  (device time of those launches, s)", BVP.kernel_ms() * 1e-3);

    for (size_t q = 0; q < m; q++) if (alive(q)) guarded(q, [&] {
        Front& f = F[q];
        const IVector X(f.XY_pt.begin(), f.XY_pt.begin() + n), Y(f.XY_pt.begin() + n, f.XY_pt.end());
        const IVector xr = f.U.containmentRatios(X), yr = f.S.containmentRatios(Y);
        f.U_new.resize(n); f.S_new.resize(n);
        for (int i = 0; i < n; i++) {
            f.U_new[i] = ia::mul(ia::mul(U_flat[i], xr[i]), interval(inflation_ratio));
            f.S_new[i] = ia::mul(ia::mul(U_flat[i], yr[i]), interval(inflation_ratio));
        }
        f.U.U_flat_global = f.U_new; f.S.U_flat_global = f.S_new;
        f.XY_nbd.assign(dimension, ia::mul(scale, initial_box));
    });
    lap("resizing");
    checkBoth(&DriverReport::conditions_resized);
    lap("manifold conditions, resized");

    for (int i = 0; i < newton_steps; i++) newtonRound(false);
    for (size_t q = 0; q < m; q++) if (alive(q)) for (int k = 0; k < dimension; k++) F[q].XY_nbd[k] = ia::mul(F[q].XY_nbd[k], interval(1.5));
    {   // the final step does not move the point or the neighbourhood (heteroclinic.cpp:246)
        std::vector<boundaryValueProblem::StepIn> in; std::vector<size_t> who;
        for (size_t q = 0; q < m; q++) if (alive(q)) { in.push_back({F[q].Mu, F[q].Ms, F[q].XY_pt, F[q].XY_nbd, F[q].T, F[q].r_u_sqr, PI[q]}); who.push_back(q); }
        const std::vector<boundaryValueProblem::StepOut> out = BVP.NewtonStepBatch(in);
        for (size_t k = 0; k < who.size(); k++) F[who[k]].success = out[k].success;
    }
    for (size_t q = 0; q < m; q++) if (alive(q)) {
        Front& f = F[q];
        bool inside = true;
        for (int i = 0; i < n; i++) {
            inside = inside && ia::subsetInterior(ia::add(f.XY_pt[i], f.XY_nbd[i]), f.U_new[i]) && ia::subsetInterior(ia::add(f.XY_pt[n + i], f.XY_nbd[n + i]), f.S_new[i]);
            rep[q].radius = std::max(rep[q].radius, std::max(ia::mag(f.XY_nbd[i]), ia::mag(f.XY_nbd[n + i])));
        }
        rep[q].inside_box = inside; rep[q].proof = f.success;
        if (!inside) code[q] = -1; else if (!f.success) code[q] = -2;
    }

    lap("Krawczyk steps and proof");
    stage_times.emplace_back("  (validated-integration launches so far)", (double)BVP.launches());
    stage_times.emplace_back("  (device time of those launches, s)", BVP.kernel_ms() * 1e-3);
    // (ii) eigenfunction error and the L_- test; end points phi(L_+)
    parallel_for(m, [&](size_t q) { if (alive(q)) guarded(q, [&] {        // per-front interval algebra: independent between fronts
        Front& f = F[q];
        f.eig = localManifold_Eig(f.U);
        f.eig.computeEigenError_minus_infty();
        rep[q].eps_minus = f.eig.eps_unscaled.hi;
        rep[q].below_Lminus = f.eig.checkConjugatePointsBelowLminus();
        const IVector X(f.XY_pt.begin(), f.XY_pt.begin() + n), Xn(f.XY_nbd.begin(), f.XY_nbd.begin() + n);
        f.start = boundaryValueProblem::getPointNbd(f.Mu, X, Xn);
        f.L_minus = interval(ia::mul(ia::mul(interval(2.0), interval(f.T)), interval(L_minus_percent)).hi);
        rep[q].L_minus = f.L_minus;
    }); });
    {
        std::vector<FlowSet> sets; std::vector<double> times; std::vector<size_t> who;
        for (size_t q = 0; q < m; q++) if (alive(q)) {
            const Front& f = F[q];
            const IVector Y(f.XY_pt.begin() + n, f.XY_pt.end()), Yn(f.XY_nbd.begin() + n, f.XY_nbd.end());
            FlowSet s = boundaryValueProblem::getPointNbd(f.Ms, Y, Yn); s.params = PI[q];
            sets.push_back(s); times.push_back(2 * f.T * (1 - L_minus_percent)); who.push_back(q);
        }
        if (!who.empty()) {
            const std::vector<boundaryValueProblem::Out> out = BVP.run(sets, times);
            for (size_t k = 0; k < who.size(); k++) {
                Front& f = F[who[k]];
                if (out[k].status != CCP_OK) { code[who[k]] = -10; continue; }
                f.endPoint = out[k].image;
                for (int i = 0; i < dimension; i++) f.endPoint[i] = ia::sub(f.endPoint[i], f.S.F.p[i]);
            }
        }
    }
    lap("eigenfunction error, L_- test, end points");
    // manifolds at L_+ for all fronts (rounds of one boundDFU launch), then (iii)-(v): one launch of the hot path
    std::vector<size_t> who;
    for (size_t q = 0; q < m; q++) if (alive(q)) who.push_back(q);
    std::vector<LPlusStage> st(who.size()); std::vector<IVector> ends(who.size());
    for (size_t k = 0; k < who.size(); k++) { st[k].dimension = dimension; st[k].manifold_subdivision = manifold_subdivision; st[k].stable = F[who[k]].S; ends[k] = F[who[k]].endPoint; }
    std::vector<char> verified;
    std::vector<localManifold_Eig> big;
    if (!who.empty()) big = constructManifoldsAtLPlusBatch(st, ends, verified);
    lap("manifolds at L_+");
    std::vector<propagateManifold> fronts; std::vector<interval> Lm; std::vector<int> thrown(who.size(), 0);
    for (size_t k = 0; k < who.size(); k++) {
        const Front& f = F[who[k]];
        fronts.emplace_back(n, PI[who[k]], f.A, f.start.p_i, f.start.Uxy, f.eig.Eu_m_Error_Final, order, stepsize);
        Lm.push_back(f.L_minus);
    }
    for (size_t k = 0; k < who.size(); k++) {
        const size_t q = who[k];
        fronts[k].below_Lminus = [&rep, q] { return rep[q].below_Lminus; };
        fronts[k].L_plus = [&st, &big, &verified, &thrown, k](const topFrame& fr, const IVector&) {
            try { return st[k].lastEuFrameWith(fr, big[k], verified[k] != 0); } catch (int e) { thrown[k] = e; return false; }   // e.g. -344 from the eigenvector validation
        };
    }
    if (!who.empty()) {
        const std::vector<int> out = propagateManifold::frameDetBatch(fronts, Lm, grid, ends, /*parallel_hooks=*/true);    // the hooks above only touch slot k
        for (size_t k = 0; k < who.size(); k++) {
            code[who[k]] = thrown[k] != 0 ? thrown[k] : out[k];
            const ccp_front_result& fr = fronts[k].frame().result();
            DriverReport& r = rep[who[k]];
            r.zero_count = fr.zero_count; r.failure_count = fr.failure_count; r.first_failure = fr.first_failure; r.series_length = fr.series_length;
        }
    }
    lap("frame count and L_+ test");
    for (size_t q = 0; q < m; q++) rep[q].result = code[q];
    if (reports) *reports = rep;
    return code;
}

}  // namespace ccp_host
