// Exercises the C++ host shim the way the reference's driver would (heteroclinic.cpp:299-334):
// build propagateManifold, call frameDet, read countZeros / getLastFrame.  Exit code 0 = ok, 77 = no GPU.
#include "../../computing-conjugate-points_b200/host/ccp_bvp.hpp"
#include "../../include/ccp_setup.h"
#include "../../computing-conjugate-points_b200/host/ccp_local_manifold.hpp"
#include <cmath>
#include <cstdio>
using namespace ccp_host;
int main() {
    if (ccp_device_count() <= 0) { std::printf("no GPU: shim compiled and linked only\n"); return 77; }
    if (ccp_init(0) != 0) { std::printf("init failed: %s\n", ccp_last_error()); return 1; }
    double params[5] = {1.0, 0.98, 0.96, 0.04, 0.02};
    ccp_front_desc d; ccp_setup_info info;
    ccp_setup_opts o{}; o.xy_nbd = -1; o.cone_L = -1; o.eig_err = -1; o.scale = -1; o.L_minus_percent = -1; o.L_minus_override = -1;
    if (ccp_setup_front(3, params, &o, &d, &info) != 0) return 2;
    const int n = 3;
    IVector par(5), p_i(6), Uxy(6); DMatrix A(6, std::vector<double>(6)); IMatrix E(3, IVector(3));
    for (int i = 0; i < 3; i++) par[i] = from_c(d.b[i]);
    for (int i = 0; i < 2; i++) par[3 + i] = from_c(d.c[i]);
    for (int i = 0; i < 6; i++) { p_i[i] = from_c(d.p_i[i]); Uxy[i] = from_c(d.Uxy[i]); for (int j = 0; j < 6; j++) A[i][j] = d.A_u[i * CCP_MAX_DIM + j]; }
    for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) E[j][k] = from_c(d.Eu_err[j * CCP_MAX_N + k]);
    propagateManifold E_u(n, par, A, p_i, Uxy, E, /*order*/ 20, /*stepsize*/ 5);
    E_u.plotting(false);
    // the two host-side proof stages are mandatory: a frameDet call with unset hooks must throw, never return a "validated" count
    bool threw = false;
    try { E_u.frameDet(interval(info.L_minus), 14, IVector(6)); } catch (const std::logic_error&) { threw = true; }
    if (!threw) { std::printf("frameDet accepted unset proof hooks\n"); return 8; }
    int unstable_e_values = E_u.frameDetUnchecked(interval(info.L_minus), /*grid*/ 14);
    std::printf("frameDet = %d, last frame U(0,0) in [%.6e, %.6e]\n", unstable_e_values, E_u.frame().getLastFrame()[0][0].lo, E_u.frame().getLastFrame()[0][0].hi);
    if (unstable_e_values != 2) return 3;
    E_u.L_plus = [](const topFrame&, const IVector&) { return false; };
    if (E_u.frameDet(interval(info.L_minus), 14, IVector(6)) != -4) return 4;
    E_u.below_Lminus = [] { return false; };
    if (E_u.frameDet(interval(info.L_minus), 14, IVector(6)) != -5) return 5;
    // boundary-value stage (boundaryValueProblem::Gxy / DGxy): forward image of the front's start set, then the backward map
    // (STABLE = true, field f_minus) of that image must contain the start point; both integrations in one launch each
    {
        boundaryValueProblem BVP(n, par, /*order*/ 20, /*stepsize*/ 5);
        FlowSet fs; fs.p_i = p_i; fs.Uxy = Uxy; fs.A_i = A; fs.stable = false;
        const double T = 4.0;
        IVector img = BVP.Gxy(fs, T);
        IMatrix DA = BVP.DGxyA(fs, T);
        FlowSet bs; bs.p_i = img; bs.Uxy = IVector(6); bs.A_i = DMatrix(6, std::vector<double>(6, 0.0)); bs.stable = true;
        for (int i = 0; i < 6; i++) bs.A_i[i][i] = 1.0;
        IVector back = BVP.Gxy(bs, T);
        for (int i = 0; i < 6; i++) {
            const double mid = 0.5 * (p_i[i].lo + p_i[i].hi);
            if (!(back[i].lo <= mid && mid <= back[i].hi) || !(back[i].hi - back[i].lo < 1e-6)) { std::printf("round trip failed in component %d\n", i); return 6; }
            if (!(DA[i][0].lo <= DA[i][0].hi)) return 7;
        }
        std::printf("Gxy round trip ok, image u_1 in [%.12e, %.12e], (D Phi A)_11 in [%.12e, %.12e]\n", img[0].lo, img[0].hi, DA[0][0].lo, DA[0][0].hi);
    }
    // boundary-value stage end to end (heteroclinic.cpp:165-281): zero-width Newton steps, Krawczyk iterations on a neighbourhood,
    // inflation by 1.5, final step with proper containment = existence + local uniqueness of the connecting orbit (given the validated
    // manifolds: cone constant L = 1.5e-5 of heteroclinic.cpp:85-86).  Every NewtonStep = 4 validated integrations in one launch.
    {
        boundaryValueProblem BVP(n, par, /*order*/ 20, /*stepsize*/ 5);
        boundaryValueProblem::Manifold Mu{A, IVector(6, interval(0.0)), interval(1.5e-5), false}, Ms{A, IVector(6, interval(0.0)), interval(1.5e-5), true};
        for (int i = 0; i < 3; i++) Ms.p[i] = interval(1.0);                       // equilibrium (1,1,1 | 0,0,0); same eigenbasis (the Hessian agrees)
        IVector XY_pt(6), XY_nbd(6, interval(0.0));
        for (int i = 0; i < 3; i++) { XY_pt[i] = interval(info.x_local[i]); XY_pt[3 + i] = interval(info.y_local[i]); }
        interval r_u_sqr(0.0);
        for (int i = 0; i < 3; i++) r_u_sqr = ia::add(r_u_sqr, ia::sqr(XY_pt[i]));
        r_u_sqr = interval(r_u_sqr.lo);
        const double T = info.T;
        auto midv = [](const IVector& v) { IVector m(v.size()); for (size_t i = 0; i < v.size(); i++) m[i] = interval(0.5 * (v[i].lo + v[i].hi)); return m; };
        // FindTime (boundaryValueProblem.cpp:118-214): from a perturbed time back to the minimiser of |Phi_T(x) - Phi_{-T}(y)|
        {
            const double Tf = BVP.FindTime(Mu, Ms, XY_pt, T + 0.02);
            std::printf("FindTime: %.6f -> %.6f (shooting time %.6f)\n", T + 0.02, Tf, T);
            if (!(std::fabs(Tf - T) < 1e-4)) return 11;
        }
        for (int it = 0; it < 6; it++) XY_pt = midv(BVP.NewtonStep(Mu, Ms, XY_pt, XY_nbd, T, r_u_sqr));
        for (int i = 0; i < 6; i++) XY_nbd[i] = interval(-1e-12, 1e-12);            // scale * initial_box
        for (int it = 0; it < 8; it++) {
            IVector out = BVP.NewtonStep(Mu, Ms, XY_pt, XY_nbd, T, r_u_sqr);
            XY_pt = midv(out);
            for (int i = 0; i < 6; i++) XY_nbd[i] = ia::sub(out[i], XY_pt[i]);
        }
        for (int i = 0; i < 6; i++) XY_nbd[i] = ia::mul(XY_nbd[i], interval(1.5));
        BVP.NewtonStep(Mu, Ms, XY_pt, XY_nbd, T, r_u_sqr);
        std::printf("BVP: T = %.6f, x = (%.6e, %.6e, %.6e), nbd radius %.3e, checkProof = %d\n", T, XY_pt[0].lo, XY_pt[1].lo, XY_pt[2].lo, XY_nbd[0].hi, (int)BVP.checkProof());
        if (!BVP.checkProof()) return 10;
    }
    // localManifold::boundDFU: the undivided box must contain the N = 15 hull, and Df at the equilibrium (diagonal in eigen-coordinates) lies inside
    {
        LocalField F; F.A = A; F.p = IVector(6, interval(0.0)); F.Ainv = IMatrix(6, IVector(6));
        // inverse of the symplectically normalised eigenbasis by Gauss-Jordan in doubles, inflated (the reference hands over CAPD's gaussInverseMatrix)
        double M[6][12];
        for (int i = 0; i < 6; i++) for (int j = 0; j < 12; j++) M[i][j] = j < 6 ? A[i][j] : (j - 6 == i ? 1.0 : 0.0);
        for (int c = 0; c < 6; c++) {
            int piv = c; for (int r = c + 1; r < 6; r++) if (std::fabs(M[r][c]) > std::fabs(M[piv][c])) piv = r;
            for (int j = 0; j < 12; j++) std::swap(M[c][j], M[piv][j]);
            const double dv = M[c][c]; for (int j = 0; j < 12; j++) M[c][j] /= dv;
            for (int r = 0; r < 6; r++) if (r != c) { const double f = M[r][c]; for (int j = 0; j < 12; j++) M[r][j] -= f * M[c][j]; }
        }
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) F.Ainv[i][j] = interval(M[i][6 + j] - 1e-12, M[i][6 + j] + 1e-12);
        IVector U(6); for (int i = 0; i < 6; i++) U[i] = i < 3 ? interval(-1e-2, 1e-2) : interval(-1e-3, 1e-3);
        IMatrix H1 = boundDFU(n, par, F, U, false, 1), H15 = boundDFU(n, par, F, U, false, 15);
        double w1 = 0, w15 = 0;
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
            if (!(H15[i][j].lo >= H1[i][j].lo - 1e-12 && H15[i][j].hi <= H1[i][j].hi + 1e-12)) { std::printf("boundDFU: subdivision is not a subset at (%d,%d)\n", i, j); return 8; }
            w1 += H1[i][j].hi - H1[i][j].lo; w15 += H15[i][j].hi - H15[i][j].lo;
        }
        if (!(w15 < w1)) return 9;
        std::printf("boundDFU ok: total width %.3e (N = 1) -> %.3e (N = 15), DF[0][0] in [%.9f, %.9f]\n", w1, w15, H15[0][0].lo, H15[0][0].hi);
    }
    ccp_shutdown();
    return 0;
}
