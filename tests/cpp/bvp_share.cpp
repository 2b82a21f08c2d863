// Krawczyk steps with the point image taken from the integration over the box (host/ccp_bvp.hpp, "shared" integrations: two per front)
// against the reference's arrangement (boundaryValueProblem.cpp:237-317: Gxy at the point and DGxy over the box, four per front), on the
// same inputs step by step.  Both operators enclose the same Newton image, so the Krawczyk images must intersect componentwise; the
// shared form may only be marginally wider, and the verdict (proper containment) must agree.  Runs against the GPU library or, linked
// with tests/cpp/oracle_dfu_backend.cpp, against the oracle twins on the CPU.  Usage: bvp_share [N_c] [N_r].  Exit 0 = ok, 77 = no GPU.
#include "../../computing-conjugate-points_b200/host/ccp_bvp.hpp"
#include "../../include/ccp_setup.h"
#include <cstdio>
#include <cstdlib>
using namespace ccp_host;
int main(int argc, char** argv) {
    if (ccp_device_count() <= 0) { std::printf("no GPU\n"); return 77; }
    if (ccp_init(0) != 0) { std::printf("init failed: %s\n", ccp_last_error()); return 1; }
    const int Nc = argc > 1 ? std::atoi(argv[1]) : 2, Nr = argc > 2 ? std::atoi(argv[2]) : 2, n = 3, D = 6;
    const size_t M = (size_t)Nc * Nr;
    const double b3[3] = {1.0, 0.975, 0.95};
    std::vector<double> P(M * 5);
    if (ccp_parameter_list(Nc, Nr, 0.02, 0.06, b3, P.data()) != 0) return 2;
    std::vector<ccp_front_desc> d(M); std::vector<ccp_setup_info> info(M);
    ccp_setup_opts o{}; o.xy_nbd = -1; o.cone_L = -1; o.eig_err = -1; o.scale = -1; o.L_minus_percent = -1; o.L_minus_override = -1;
    if (ccp_setup_sweep(n, P.data(), M, &o, d.data(), info.data(), 0) != 0) return 3;
    boundaryValueProblem shared(n, IVector(5, interval(0.0)), 20, 3), four(n, IVector(5, interval(0.0)), 20, 3);
    shared.share_integrations(true); four.share_integrations(false);
    shared.pipeline_chunks(2, 1); four.pipeline_chunks(1);             // the shared arrangement also goes through the chunked host pipeline
    std::vector<boundaryValueProblem::StepIn> in(M);
    for (size_t f = 0; f < M; f++) {
        boundaryValueProblem::StepIn& q = in[f];
        DMatrix A(D, std::vector<double>(D));
        for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) A[i][j] = d[f].A_u[i * CCP_MAX_DIM + j];
        q.unstable = boundaryValueProblem::Manifold{A, IVector(D, interval(0.0)), interval(1.5e-5), false};
        q.stable = boundaryValueProblem::Manifold{A, IVector(D, interval(0.0)), interval(1.5e-5), true};
        for (int i = 0; i < n; i++) q.stable.p[i] = interval(1.0);
        q.params.resize(5); for (int i = 0; i < 5; i++) q.params[i] = interval(P[f * 5 + i]);
        q.XY_pt.resize(D); q.XY_nbd.assign(D, interval(0.0));
        for (int i = 0; i < n; i++) { q.XY_pt[i] = interval(info[f].x_local[i]); q.XY_pt[n + i] = interval(info[f].y_local[i]); }
        interval r2(0.0); for (int i = 0; i < n; i++) r2 = ia::add(r2, ia::sqr(q.XY_pt[i]));
        q.r_u_sqr = interval(r2.lo); q.T = info[f].T;
    }
    auto midv = [](const IVector& v) { IVector m(v.size()); for (size_t i = 0; i < v.size(); i++) m[i] = interval(0.5 * (v[i].lo + v[i].hi)); return m; };
    for (int it = 0; it < 3; it++) { auto r = shared.NewtonStepBatch(in); for (size_t f = 0; f < M; f++) if (!r[f].kraw_image.empty()) in[f].XY_pt = midv(r[f].kraw_image); }
    const int l_thin = shared.launches();
    for (size_t f = 0; f < M; f++) in[f].XY_nbd.assign(D, interval(-1e-12, 1e-12));
    double worst = 0; int bad = 0; size_t ok1 = 0, ok0 = 0;
    for (int it = 0; it < 3; it++) {
        if (it == 2) for (size_t f = 0; f < M; f++) for (int i = 0; i < D; i++) in[f].XY_nbd[i] = ia::mul(in[f].XY_nbd[i], interval(1.5));
        auto r1 = shared.NewtonStepBatch(in); auto r0 = four.NewtonStepBatch(in);
        ok1 = ok0 = 0;
        for (size_t f = 0; f < M; f++) {
            ok1 += r1[f].success; ok0 += r0[f].success;
            if (r1[f].success != r0[f].success) bad++;
            for (int i = 0; i < D; i++) {
                const interval a = r1[f].kraw_image[i], b = r0[f].kraw_image[i];
                if (a.hi < b.lo || b.hi < a.lo) { std::printf("front %zu component %d: disjoint Krawczyk images\n", f, i); bad++; }
                const double wa = a.hi - a.lo, wb = b.hi - b.lo;
                if (wb > 0) worst = std::max(worst, wa / wb);
            }
        }
        for (size_t f = 0; f < M; f++) { in[f].XY_pt = midv(r1[f].kraw_image); for (int i = 0; i < D; i++) in[f].XY_nbd[i] = ia::sub(r1[f].kraw_image[i], in[f].XY_pt[i]); }
    }
    std::printf("bvp_share: %zu fronts, verified %zu (shared) / %zu (four integrations), widest shared image %.3f x the four-integration image, %d disagreements; thin steps used %d launches\n",
                M, ok1, ok0, worst, bad, l_thin);
    ccp_shutdown();
    return (bad == 0 && worst <= 1.25 && ok1 == ok0) ? 0 : 4;
}
