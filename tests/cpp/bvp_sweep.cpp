// Boundary-value stage for a parameter sweep: the driver's Newton / Krawczyk loop (heteroclinic.cpp:186-281) for M
// fronts of Construct_ParameterList at once -- every NewtonStepBatch is ONE launch of 4*M validated integrations.  Prints the number of
// fronts whose connecting orbit is verified (given the cone constant of the manifold validation) and the time per batched Krawczyk step.
// Usage: bvp_sweep [N_c] [N_r] [step_size] [share] [chunks]   (default 24 x 4 = the reference's own 96-point sweep, first rung of the step ladder 2^-3,
// share = 1: Krawczyk steps take the point image from the integration over the box -- 2 integrations per front instead of 4).  Exit 0 = all verified, 77 = no GPU.
#include "../../computing-conjugate-points_b200/host/ccp_bvp.hpp"
#include "../../include/ccp_setup.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
using namespace ccp_host;
int main(int argc, char** argv) {
    if (ccp_device_count() <= 0) { std::printf("no GPU\n"); return 77; }
    if (ccp_init(0) != 0) { std::printf("init failed: %s\n", ccp_last_error()); return 1; }
    const int Nc = argc > 1 ? std::atoi(argv[1]) : 24, Nr = argc > 2 ? std::atoi(argv[2]) : 4, n = 3, D = 6;
    const size_t M = (size_t)Nc * Nr;
    const double b3[3] = {1.0, 0.975, 0.95};
    std::vector<double> P(M * 5);
    if (ccp_parameter_list(Nc, Nr, 0.02, 0.06, b3, P.data()) != 0) return 2;
    std::vector<ccp_front_desc> d(M); std::vector<ccp_setup_info> info(M);
    ccp_setup_opts o{}; o.xy_nbd = -1; o.cone_L = -1; o.eig_err = -1; o.scale = -1; o.L_minus_percent = -1; o.L_minus_override = -1;
    if (ccp_setup_sweep(n, P.data(), M, &o, d.data(), info.data(), 0) != 0) return 3;
    const int step_size = argc > 3 ? std::atoi(argv[3]) : 3;                    // fixed step 2^-step_size of the validated integrations
    boundaryValueProblem BVP(n, IVector(5, interval(0.0)), 20, step_size);
    const bool share = argc > 4 ? std::atoi(argv[4]) != 0 : true;
    BVP.share_integrations(share);
    const int chunks = argc > 5 ? std::atoi(argv[5]) : 2;                       // host pipelines per batched step (1 = off)
    BVP.pipeline_chunks(chunks);
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
    double ms_step = 0, ms_box = 0, dev_box = 0; int timed = 0, timed_box = 0; bool box_phase = false;
    auto step = [&]() {
        const double k0 = BVP.kernel_ms();
        auto t0 = std::chrono::steady_clock::now();
        std::vector<boundaryValueProblem::StepOut> r = BVP.NewtonStepBatch(in);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        ms_step += ms; timed++;
        if (box_phase) { ms_box += ms; dev_box += BVP.kernel_ms() - k0; timed_box++; }
        return r;
    };
    { BVP.NewtonStepBatch(in); }                                                // untimed warm-up (module load, first allocation of the staging buffers)
    const double warm_ms = BVP.kernel_ms(); const int warm_launches = BVP.launches();
    for (int it = 0; it < 4; it++) { auto r = step(); for (size_t f = 0; f < M; f++) if (!r[f].kraw_image.empty()) in[f].XY_pt = midv(r[f].kraw_image); }
    for (size_t f = 0; f < M; f++) in[f].XY_nbd.assign(D, interval(-1e-12, 1e-12));
    box_phase = true;
    for (int it = 0; it < 5; it++) {
        auto r = step();
        for (size_t f = 0; f < M; f++) { in[f].XY_pt = midv(r[f].kraw_image); for (int i = 0; i < D; i++) in[f].XY_nbd[i] = ia::sub(r[f].kraw_image[i], in[f].XY_pt[i]); }
    }
    for (size_t f = 0; f < M; f++) for (int i = 0; i < D; i++) in[f].XY_nbd[i] = ia::mul(in[f].XY_nbd[i], interval(1.5));
    auto r = step();
    size_t ok = 0; double rmax = 0;
    for (size_t f = 0; f < M; f++) { ok += r[f].success ? 1 : 0; for (int i = 0; i < D; i++) rmax = std::max(rmax, in[f].XY_nbd[i].hi); }
    std::printf("BVP sweep: %zu fronts, %zu connecting orbits verified, largest neighbourhood radius %.3e; %d batched Krawczyk steps, %.1f ms per step (up to %zu validated integrations each, host algebra included; %.1f ms of it on the device in %.1f launches) = %.0f front-steps/s\n",
                M, ok, rmax, timed, ms_step / timed, 4 * M, (BVP.kernel_ms() - warm_ms) / timed, (double)(BVP.launches() - warm_launches) / timed, M / (ms_step / timed) * 1e3);
    std::printf("  Krawczyk steps with a neighbourhood (XY_nbd != 0, %s): %d steps, %.1f ms per step, %.1f ms of it on the device\n",
                share ? "2 integrations per front: point image from the integration over the box" : "4 integrations per front", timed_box, ms_box / timed_box, dev_box / timed_box);
    std::printf("  host pipelines per step: %d\n", chunks);
    ccp_shutdown();
    return ok == M ? 0 : 4;
}
