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
// Integration parameters are the engine's (order; step ladder 2^-step_size .. 2^-finest_step, see the constructor); T is a point (the
// reference passes a thin interval).
#pragma once
#include "ccp_frame_det.hpp"
#include <algorithm>
#include <cmath>
#include <mutex>
#include <string>
#include <thread>

namespace ccp_host {

// Rigorous host interval operations for the small amount of interval algebra of NewtonStep: every IEEE result is moved outward by
// one ulp (nextafter), which encloses the exact result of a correctly rounded +, -, *, sqrt whatever the rounding mode.
namespace ia {
inline double dn(double x) { return std::nextafter(x, -INFINITY); }
inline double up(double x) { return std::nextafter(x, INFINITY); }
inline interval add(const interval& a, const interval& b) { return interval(dn(a.lo + b.lo), up(a.hi + b.hi)); }
inline interval sub(const interval& a, const interval& b) { return interval(dn(a.lo - b.hi), up(a.hi - b.lo)); }
inline interval neg(const interval& a) { return interval(-a.hi, -a.lo); }
inline interval mul(const interval& a, const interval& b) {
    const double c[4] = {a.lo * b.lo, a.lo * b.hi, a.hi * b.lo, a.hi * b.hi};
    return interval(dn(std::min(std::min(c[0], c[1]), std::min(c[2], c[3]))), up(std::max(std::max(c[0], c[1]), std::max(c[2], c[3]))));
}
inline interval sqr(const interval& a) {
    if (a.lo >= 0) return interval(dn(a.lo * a.lo), up(a.hi * a.hi));
    if (a.hi <= 0) return interval(dn(a.hi * a.hi), up(a.lo * a.lo));
    return interval(0.0, up(std::max(a.lo * a.lo, a.hi * a.hi)));
}
inline interval sqrt(const interval& a) { return interval(a.lo > 0 ? std::max(0.0, dn(std::sqrt(a.lo))) : 0.0, up(std::sqrt(std::max(a.hi, 0.0)))); }
}

struct FlowSet {                 // what getPointNbd / (*pF).A hand to C0Rect2Set / C1Rect2Set
    IVector p_i, Uxy;            // 2n each
    DMatrix A_i;                 // 2n x 2n point matrix
    bool stable = false;         // true: backward flow from the stable manifold (f_minus)
    IVector params;              // optional: parameters of this set (sweeps over parameter values); empty = the solver's
};

class boundaryValueProblem {
  public:
    // `step_size`: the validated integrations start with the step 2^-step_size.  The reference's solvers of this stage are adaptive
    // (ITaylor without setStep, boundaryValueProblem.cpp:24-38: CAPD picks the step); here the step ladder is explicit: every integration
    // of a launch first runs with the COARSE step (default 2^-3: with order 20 the truncation bound is still ~1e-22 and the a-priori
    // enclosure of a step holds along the whole front) and the ones whose enclosure check fails (per-front status, CCP_ST_ENCLOSURE)
    // are re-launched with the next finer step, down to 2^-finest_step.  The result of any rung is a valid enclosure.
    boundaryValueProblem(int n, const IVector& params, int order, int step_size = 3, int finest_step = 6)
        : n_(n), params_(params), order_(order), step_(step_size), finest_(finest_step < step_size ? step_size : finest_step) {
        if (n < 2 || n > CCP_MAX_N) throw std::invalid_argument("n out of range");
    }
    void share_integrations(bool on) { share_ = on; }     // Krawczyk phase: point image from the integration over the box (default on)
    int launches() const { return launches_; }            // launches of validated integrations so far (all rungs)
    double kernel_ms() const { return kernel_ms_; }       // device time of those launches (slowest device per launch)
    // enclosure of the time-T image of the set (Gxy)
    IVector Gxy(const FlowSet& s, double T) { return run({s}, {T})[0].image; }
    // enclosure of [D Phi_T] * A_i over the set (DGxy before the factor DW)
    IMatrix DGxyA(const FlowSet& s, double T) { return run({s}, {T})[0].deriv_A; }

    struct Out { int status; IVector image, centre; IMatrix deriv_A; };      // centre: image of p_i alone (rho = 0), see ccp.h last_frame column n
    // all integrations of a Newton step (or of a whole sweep) in ONE launch
    std::vector<Out> run(const std::vector<FlowSet>& sets, const std::vector<double>& T) {
        const int D = 2 * n_;
        std::vector<ccp_front_desc> d(sets.size());
        std::vector<ccp_front_result> r(sets.size());
        parallel_for(sets.size(), [&](size_t q) {
            const FlowSet& s = sets[q];
            ccp_front_desc& e = d[q];
            e = ccp_front_desc{};
            e.n = n_; e.order = order_; e.step_log2 = step_; e.grid = 0; e.L_minus = T[q];     // grid 0 = flow map only
            const IVector& par = s.params.empty() ? params_ : s.params;
            for (int i = 0; i < n_; i++) e.b[i] = to_c(par[i]);
            for (int i = 0; i < n_ - 1; i++) e.c[i] = to_c(par[n_ + i]);
            for (int i = 0; i < D; i++) {
                const double sg = (s.stable && i >= n_) ? -1.0 : 1.0;                       // S = diag(I, -I)
                e.p_i[i] = sg > 0 ? to_c(s.p_i[i]) : to_c(interval(-s.p_i[i].hi, -s.p_i[i].lo));
                e.Uxy[i] = to_c(s.Uxy[i]);
                for (int j = 0; j < D; j++) e.A_u[i * CCP_MAX_DIM + j] = sg * s.A_i[i][j];
            }
        });
        {
            std::lock_guard<std::mutex> lk(dev_mu_);
            if (ccp_frame_count_batch_multi(d.data(), d.size(), r.data()) != 0) throw std::runtime_error(std::string("ccp: ") + ccp_last_error());
            launches_++; kernel_ms_ += ccp_last_kernel_ms();
        }
        // step ladder: integrations whose a-priori enclosure failed with the coarse step are re-run with the next finer one
        for (int st = step_ + 1; st <= finest_; st++) {
            std::vector<size_t> redo;
            for (size_t q = 0; q < d.size(); q++) if (r[q].status == CCP_ST_ENCLOSURE) redo.push_back(q);
            if (redo.empty()) break;
            std::vector<ccp_front_desc> d2(redo.size()); std::vector<ccp_front_result> r2(redo.size());
            for (size_t k = 0; k < redo.size(); k++) { d2[k] = d[redo[k]]; d2[k].step_log2 = st; }
            {
                std::lock_guard<std::mutex> lk(dev_mu_);
                if (ccp_frame_count_batch_multi(d2.data(), d2.size(), r2.data()) != 0) throw std::runtime_error(std::string("ccp: ") + ccp_last_error());
                launches_++; kernel_ms_ += ccp_last_kernel_ms();
            }
            for (size_t k = 0; k < redo.size(); k++) r[redo[k]] = r2[k];
        }
        std::vector<Out> out(sets.size());
        parallel_for(sets.size(), [&](size_t q) {
            Out& o = out[q];
            o.status = r[q].status;
            o.image.resize(D); o.centre.resize(D); o.deriv_A.assign(D, IVector(D));
            for (int i = 0; i < D; i++) {
                const bool flip = sets[q].stable && i >= n_;
                auto m = [&](const ccp_interval& a) { return flip ? interval(-a.hi, -a.lo) : from_c(a); };
                o.image[i] = m(r[q].last_frame[i * (CCP_MAX_N + 2) + n_ + 1]);
                o.centre[i] = m(r[q].last_frame[i * (CCP_MAX_N + 2) + n_]);
                for (int j = 0; j < D; j++) o.deriv_A[i][j] = m(r[q].flow_P[i * CCP_MAX_DIM + j]);
            }
        });
        return out;
    }

    // ---------------- the Krawczyk / Newton step of the boundary-value stage ----------------
    // What boundaryValueProblem reads from its two localManifold objects (localManifold.h:24-36, localVField.h:17-19).
    struct Manifold { DMatrix A; IVector p; interval L; bool stable; };

    // localManifold::getPointNbd (localManifold.cpp:73-102) with constructU (:4-41) and getPoint (:43-71)
    static FlowSet getPointNbd(const Manifold& m, const IVector& pt, const IVector& nbd) {
        const int n = (int)pt.size(), D = 2 * n;
        interval norm(0.0);
        for (int i = 0; i < n; i++) norm = ia::add(norm, ia::sqr(ia::add(pt[i], nbd[i])));
        norm = ia::sqrt(norm);
        const interval error = ia::mul(ia::mul(m.L, interval(-1.0, 1.0)), norm);        // +- L ||U_flat||
        FlowSet s; s.A_i = m.A; s.stable = m.stable; s.p_i.assign(D, interval(0.0)); s.Uxy.assign(D, interval(0.0));
        IVector local(D, interval(0.0));
        for (int i = 0; i < n; i++) local[m.stable ? n + i : i] = pt[i];
        for (int i = 0; i < D; i++) {
            const bool on = m.stable ? i >= n : i < n;
            const interval U = on ? ia::add(pt[m.stable ? i - n : i], nbd[m.stable ? i - n : i]) : error;   // constructU(XY_pt + XY_nbd)
            s.Uxy[i] = ia::sub(U, local[i]);                                                                  // translated back to enclose zero
            interval a = m.p[i];
            for (int j = 0; j < D; j++) a = ia::add(a, ia::mul(interval(m.A[i][j]), local[j]));              // p + A * local_pt
            s.p_i[i] = a;
        }
        return s;
    }
    // localManifold::constructDW (localManifold.cpp:134-175): chart derivative, identity on the manifold coordinates, +-L elsewhere
    static IMatrix DW(const Manifold& m, int n) {
        IMatrix W(2 * n, IVector(n));
        const interval error = ia::mul(m.L, interval(-1.0, 1.0));
        for (int i = 0; i < 2 * n; i++) for (int j = 0; j < n; j++) {
            const bool on = m.stable ? i >= n : i < n;
            W[i][j] = on ? interval((m.stable ? i - n : i) == j ? 1.0 : 0.0) : error;
        }
        return W;
    }

    bool SUCCESS = false;
    // One Krawczyk step of one front: everything NewtonStep reads (the two manifolds, the point and its neighbourhood, T, r_u^2).
    struct StepIn { Manifold unstable, stable; IVector XY_pt, XY_nbd; double T; interval r_u_sqr; IVector params; };
    struct StepOut { IVector kraw_image; bool success; };

    // boundaryValueProblem::NewtonStep (boundaryValueProblem.cpp:237-317) for many fronts (a parameter sweep): the image of
    // XY_pt + XY_nbd under the interval Krawczyk operator of G(x,y) = Phi_T(x) - Phi_{-T}(y) (last row: |x|^2 - r_u^2).
    // The validated integrations of every front -- 4 * in.size(), 2 * in.size() while XY_nbd = 0 -- are ONE launch (per rung of the step ladder).
    // Large batches are cut into chunks that run as concurrent host pipelines (marshalling -> launch -> Krawczyk algebra): the engine
    // serialises the launches, so the host algebra of one chunk overlaps the integrations of the next.  A chunk keeps >= ~1,000
    // integrations (one wave of the front kernel), otherwise the launches themselves would get slower.  pipeline_chunks(1) = off.
    void pipeline_chunks(int k, size_t min_fronts = 512) { chunks_ = k < 1 ? 1 : k; chunk_min_ = min_fronts < 1 ? 1 : min_fronts; }
    std::vector<StepOut> NewtonStepBatch(const std::vector<StepIn>& in) {
        size_t K = std::min<size_t>((size_t)chunks_, in.size() / chunk_min_);
        if (K <= 1) return NewtonStepChunk(in.data(), in.size());
        std::vector<std::vector<StepOut>> part(K);
        std::vector<std::thread> th;
        std::vector<std::string> err(K);
        const size_t per = (in.size() + K - 1) / K;
        for (size_t c = 0; c < K; c++) {
            const size_t lo = c * per, hi = std::min(in.size(), lo + per);
            th.emplace_back([&, c, lo, hi] { try { part[c] = NewtonStepChunk(in.data() + lo, hi - lo); } catch (const std::exception& e) { err[c] = e.what(); if (err[c].empty()) err[c] = "error"; } });
        }
        for (auto& t : th) t.join();
        for (size_t c = 0; c < K; c++) if (!err[c].empty()) throw std::runtime_error(err[c]);
        std::vector<StepOut> res; res.reserve(in.size());
        for (size_t c = 0; c < K; c++) for (auto& r : part[c]) res.push_back(std::move(r));
        return res;
    }
    std::vector<StepOut> NewtonStepChunk(const StepIn* inp, size_t in_size) {
        struct View { const StepIn* p; size_t n; size_t size() const { return n; } const StepIn& operator[](size_t i) const { return p[i]; } } in{inp, in_size};
        const int n = n_, D = 2 * n;
        const IVector zero(n, interval(0.0));
        // The operator needs G at the POINT (C^0 images of XY_pt, sets 0 and 1) and DG over the BOX (C^1 maps of XY_pt + XY_nbd, sets 2
        // and 3).  In the Newton phase of the driver (heteroclinic.cpp:186-191: 20 steps with XY_nbd = 0) box and point coincide, and
        // since one integration delivers image and derivative alike only two of the four are launched.
        // Krawczyk phase (XY_nbd != 0): the point set p_i + A_i U_pt and the box set p_i + A_i U_box share p_i, and U_pt is a sub-box of
        // U_box (the neighbourhood and the cone error +-L|x| both only grow with XY_nbd).  The integration over the box returns the image of
        // p_i alone (Out::centre) and [D Phi_T(box set)] A_i (Out::deriv_A), hence by the mean-value form along the segment p_i .. p_i + A_i rho
        //     Phi_T(p_i + A_i rho)  in  centre + deriv_A rho      for every rho in U_pt,
        // a valid enclosure of the point image G(x) -- the same two integrations per front as in the Newton phase instead of four
        // (share_integrations(false) restores the reference's four; a front whose sets do not nest falls back to four on its own).
        std::vector<size_t> slot(4 * in.size()), first(in.size() + 1, 0);
        std::vector<char> thin(in.size(), 1), shared(in.size(), 0);
        std::vector<FlowSet> four(4 * in.size());
        parallel_for(in.size(), [&](size_t fi) {
            const StepIn& q = in[fi];
            for (const interval& e : q.XY_nbd) thin[fi] = thin[fi] && e.lo == 0.0 && e.hi == 0.0;
            IVector X_pt(q.XY_pt.begin(), q.XY_pt.begin() + n), Y_pt(q.XY_pt.begin() + n, q.XY_pt.end());
            IVector X_nbd(q.XY_nbd.begin(), q.XY_nbd.begin() + n), Y_nbd(q.XY_nbd.begin() + n, q.XY_nbd.end());
            bool nest = share_ && !thin[fi];
            for (int k = 0; k < 4; k++) {
                if (thin[fi] && k >= 2) continue;
                FlowSet fs = getPointNbd(k % 2 ? q.stable : q.unstable, k % 2 ? Y_pt : X_pt, k < 2 ? zero : (k % 2 ? Y_nbd : X_nbd));
                fs.params = q.params;
                four[4 * fi + k] = std::move(fs);
            }
            for (int k = 0; k < 2 && nest; k++) {
                const FlowSet &pt = four[4 * fi + k], &bx = four[4 * fi + 2 + k];
                for (int i = 0; i < D && nest; i++)
                    nest = pt.p_i[i].lo == bx.p_i[i].lo && pt.p_i[i].hi == bx.p_i[i].hi && bx.Uxy[i].lo <= pt.Uxy[i].lo && pt.Uxy[i].hi <= bx.Uxy[i].hi &&
                           bx.Uxy[i].lo <= 0.0 && 0.0 <= bx.Uxy[i].hi;
            }
            shared[fi] = nest ? 1 : 0;
        });
        for (size_t fi = 0; fi < in.size(); fi++) first[fi + 1] = first[fi] + ((thin[fi] || shared[fi]) ? 2 : 4);
        std::vector<FlowSet> sets(first[in.size()]); std::vector<double> times(first[in.size()]);
        for (size_t fi = 0; fi < in.size(); fi++) {
            if (thin[fi]) for (int k = 0; k < 4; k++) slot[4 * fi + k] = first[fi] + k % 2;
            else if (shared[fi]) for (int k = 0; k < 4; k++) slot[4 * fi + k] = first[fi] + k % 2;       // the box sets only
            else for (int k = 0; k < 4; k++) slot[4 * fi + k] = first[fi] + k;
            const int nk = (thin[fi] || shared[fi]) ? 2 : 4;
            for (int k = 0; k < nk; k++) times[first[fi] + k] = in[fi].T;
        }
        parallel_for(in.size(), [&](size_t fi) {                     // the sets that are launched move out of `four`; the point sets of shared fronts stay (their Uxy is U_pt)
            const int nk = (thin[fi] || shared[fi]) ? 2 : 4;
            for (int k = 0; k < nk; k++) sets[first[fi] + k] = std::move(four[4 * fi + (shared[fi] ? 2 + k : k)]);
        });
        const std::vector<Out> all = run(sets, times);
        std::vector<StepOut> res(in.size());
        parallel_for(in.size(), [&](size_t f) {
            const StepIn& q = in[f];
            const Out o[4] = {all[slot[4 * f]], all[slot[4 * f + 1]], all[slot[4 * f + 2]], all[slot[4 * f + 3]]};
            res[f].success = false; res[f].kraw_image.assign(D, interval(0.0));
            bool okflow = true; for (int k = 0; k < 4; k++) okflow = okflow && o[k].status == CCP_OK;
            if (!okflow) return;                                     // like a CAPD exception inside one parameter value: this front fails, the sweep goes on
            IVector X_pt(q.XY_pt.begin(), q.XY_pt.begin() + n), X_nbd(q.XY_nbd.begin(), q.XY_nbd.begin() + n);
            IVector G_x = o[0].image, G_y = o[1].image;
            if (shared[f]) {                                         // point images from the integrations over the boxes (mean-value form, see above)
                for (int s2 = 0; s2 < 2; s2++) {
                    const Out& ob = o[2 + s2]; const IVector& rho = four[4 * f + s2].Uxy;
                    IVector& Gp = s2 ? G_y : G_x;
                    for (int i = 0; i < D; i++) {
                        interval a = ob.centre[i];
                        for (int k = 0; k < D; k++) a = ia::add(a, ia::mul(ob.deriv_A[i][k], rho[k]));
                        Gp[i] = a;
                    }
                }
            }
            IVector G(D);
            for (int i = 0; i < D; i++) G[i] = ia::sub(G_x[i], G_y[i]);
            interval x_radius_sqr(0.0);
            for (int i = 0; i < n; i++) x_radius_sqr = ia::add(x_radius_sqr, ia::sqr(X_pt[i]));
            G[D - 1] = ia::sub(x_radius_sqr, q.r_u_sqr);
            // DG = [ DPhi_T A_u DW_u | -DPhi_{-T} A_s DW_s ;  2 (X_pt + X_nbd) | 0 ]      (DGxy :108-111, DG_combine :320-353)
            const IMatrix Wu = DW(q.unstable, n), Ws = DW(q.stable, n);
            IMatrix DG(D, IVector(D, interval(0.0)));
            for (int i = 0; i < D; i++) for (int j = 0; j < n; j++) {
                interval a(0.0), b(0.0);
                for (int k = 0; k < D; k++) { a = ia::add(a, ia::mul(o[2].deriv_A[i][k], Wu[k][j])); b = ia::add(b, ia::mul(o[3].deriv_A[i][k], Ws[k][j])); }
                DG[i][j] = a; DG[i][n + j] = ia::neg(b);
            }
            for (int j = 0; j < D; j++) DG[D - 1][j] = j < n ? ia::mul(interval(2.0), ia::add(X_pt[j], X_nbd[j])) : interval(0.0);
            // fixed approximate inverse of mid(DG) (the reference: midMatrix(krawczykInverse(midMatrix(DG))) -- any approximation is valid)
            std::vector<std::vector<double>> M(D, std::vector<double>(2 * D, 0.0));
            for (int i = 0; i < D; i++) { for (int j = 0; j < D; j++) M[i][j] = 0.5 * (DG[i][j].lo + DG[i][j].hi); M[i][D + i] = 1.0; }
            bool singular = false;
            for (int c = 0; c < D && !singular; c++) {
                int piv = c; for (int r = c + 1; r < D; r++) if (std::fabs(M[r][c]) > std::fabs(M[piv][c])) piv = r;
                std::swap(M[c], M[piv]);
                const double dv = M[c][c]; if (dv == 0.0 || !std::isfinite(dv)) { singular = true; break; }
                for (int j = 0; j < 2 * D; j++) M[c][j] /= dv;
                for (int r = 0; r < D; r++) if (r != c) { const double fct = M[r][c]; for (int j = 0; j < 2 * D; j++) M[r][j] -= fct * M[c][j]; }
            }
            if (singular) return;
            // new_Nbd = -C G + (I - C DG) XY_nbd
            IVector new_Nbd(D);
            for (int i = 0; i < D; i++) {
                interval a(0.0);
                for (int k = 0; k < D; k++) a = ia::sub(a, ia::mul(interval(M[i][D + k]), G[k]));
                for (int j = 0; j < D; j++) {
                    interval e(i == j ? 1.0 : 0.0);
                    for (int k = 0; k < D; k++) e = ia::sub(e, ia::mul(interval(M[i][D + k]), DG[k][j]));
                    a = ia::add(a, ia::mul(e, q.XY_nbd[j]));
                }
                new_Nbd[i] = a; res[f].kraw_image[i] = ia::add(q.XY_pt[i], a);
            }
            // same energy level: the discarded velocity component must have a definite sign (:303-313); then proper containment
            const interval gy = G_y[D - 1];
            bool ok = gy.lo > 0 || gy.hi < 0;
            for (int i = 0; i < D && ok; i++) ok = q.XY_nbd[i].lo < new_Nbd[i].lo && new_Nbd[i].hi < q.XY_nbd[i].hi;     // subsetInterior
            res[f].success = ok;
        });
        return res;
    }
    // the reference's single-front signature
    IVector NewtonStep(const Manifold& unstable, const Manifold& stable, const IVector& XY_pt, const IVector& XY_nbd, double T, interval r_u_sqr) {
        std::vector<StepOut> r = NewtonStepBatch({StepIn{unstable, stable, XY_pt, XY_nbd, T, r_u_sqr, IVector()}});
        SUCCESS = r[0].success;
        return r[0].kraw_image;
    }
    bool checkProof() const { return SUCCESS; }          // boundaryValueProblem::checkProof

    // boundaryValueProblem::FindTime (boundaryValueProblem.cpp:118-214): for fixed points x, y on the two manifolds, 10 Newton steps
    // on T for F(T) = <Gx - Gy, f(Gx) + f(Gy)> = d/dT |Phi_T(x) - Phi_{-T}(y)|^2 / 2; like the reference only midpoints enter the
    // update, so f and Df are evaluated in doubles at the midpoints of the validated images (two integrations per step, one launch).
    double FindTime(const Manifold& unstable, const Manifold& stable, const IVector& XY_pt, double T) {
        return FindTimeBatch({TimeIn{unstable, stable, XY_pt, T, IVector()}})[0];
    }
    // the same for many fronts (a parameter sweep): every Newton step on T is ONE launch of two integrations per front still iterating
    struct TimeIn { Manifold unstable, stable; IVector XY_pt; double T; IVector params; };
    std::vector<double> FindTimeBatch(const std::vector<TimeIn>& in) {
        const int n = n_, D = 2 * n;
        const size_t m = in.size();
        const IVector zero(n, interval(0.0));
        auto mid = [](const interval& a) { return 0.5 * (a.lo + a.hi); };
        auto field = [&](const IVector& par, const std::vector<double>& z, std::vector<double>& f, std::vector<std::vector<double>>& Df) {
            std::vector<double> b(n), c(n, 0.0);
            for (int i = 0; i < n; i++) b[i] = mid(par[i]);
            for (int i = 0; i < n - 1; i++) c[i] = mid(par[n + i]);
            f.assign(D, 0.0); Df.assign(D, std::vector<double>(D, 0.0));
            std::vector<double> g(n), w(n);
            for (int i = 0; i < n; i++) { g[i] = z[i] * (1.0 - z[i]); w[i] = z[i] - 0.5; f[i] = z[n + i]; Df[i][n + i] = 1.0; }
            for (int i = 0; i < n; i++) {
                double F = -b[i] * g[i] * w[i], Mii = -b[i] * (3.0 * g[i] - 0.5);
                if (i > 0)     { F -= c[i - 1] * g[i - 1] * w[i]; Mii -= c[i - 1] * g[i - 1]; Df[n + i][i - 1] = 2.0 * c[i - 1] * w[i] * w[i - 1]; }
                if (i < n - 1) { F -= c[i] * g[i + 1] * w[i];     Mii -= c[i] * g[i + 1];     Df[n + i][i + 1] = 2.0 * c[i] * w[i] * w[i + 1]; }
                f[n + i] = F; Df[n + i][i] = Mii;
            }
        };
        std::vector<double> T(m), original_dist(m, 0.0), dist0(m, 0.0);
        std::vector<char> active(m, 1), failed(m, 0);
        for (size_t q = 0; q < m; q++) T[q] = in[q].T;
        for (int it = 0; it < 10; it++) {
            std::vector<FlowSet> sets; std::vector<double> times; std::vector<size_t> who;
            for (size_t q = 0; q < m; q++) {
                if (!active[q]) continue;
                const IVector X_pt(in[q].XY_pt.begin(), in[q].XY_pt.begin() + n), Y_pt(in[q].XY_pt.begin() + n, in[q].XY_pt.end());
                FlowSet fx = getPointNbd(in[q].unstable, X_pt, zero), fy = getPointNbd(in[q].stable, Y_pt, zero);
                fx.params = in[q].params; fy.params = in[q].params;
                sets.push_back(fx); sets.push_back(fy); times.push_back(T[q]); times.push_back(T[q]); who.push_back(q);
            }
            if (who.empty()) break;
            const std::vector<Out> o = run(sets, times);
            for (size_t k = 0; k < who.size(); k++) {
                const size_t q = who[k];
                const Out &ox = o[2 * k], &oy = o[2 * k + 1];
                if (ox.status != CCP_OK || oy.status != CCP_OK) { active[q] = 0; failed[q] = 1; continue; }
                const IVector& par = in[q].params.empty() ? params_ : in[q].params;
                std::vector<double> gx(D), gy(D), fx, fy; std::vector<std::vector<double>> Dx, Dy;
                for (int i = 0; i < D; i++) { gx[i] = mid(ox.image[i]); gy[i] = mid(oy.image[i]); }
                field(par, gx, fx, Dx); field(par, gy, fy, Dy);
                double d0 = 0, d1 = 0, d2 = 0;
                for (int i = 0; i < D; i++) {
                    // the y-trajectory runs backwards: d/dT Phi_{-T}(y) = -f, d^2/dT^2 = +Df f
                    const double dv = gx[i] - gy[i], dd1 = fx[i] + fy[i];
                    double ax = 0, ay = 0; for (int kk = 0; kk < D; kk++) { ax += Dx[i][kk] * fx[kk]; ay += Dy[i][kk] * fy[kk]; }
                    const double dd2 = ax - ay;
                    d0 += dv * dv; d1 += dv * dd1; d2 += dd1 * dd1 + dv * dd2;
                }
                dist0[q] = d0;
                if (it == 0) original_dist[q] = d0;
                if (d2 == 0.0 || !std::isfinite(d1 / d2)) { active[q] = 0; continue; }
                T[q] = T[q] - d1 / d2;
            }
        }
        for (size_t q = 0; q < m; q++) if (failed[q] || original_dist[q] < dist0[q]) T[q] = in[q].T;
        return T;
    }

  private:
    int n_; IVector params_; int order_, step_, finest_; int launches_ = 0; double kernel_ms_ = 0.0; bool share_ = true; int chunks_ = 2; size_t chunk_min_ = 512; std::mutex dev_mu_;
};

} // namespace ccp_host
