// frameDetBatch over every GPU of the box from ONE C++ host thread (ccp_init_multi + ccp_frame_count_batch_multi): the replacement
// of the reference's serial sweep loop (heteroclinic.cpp:449-471) that uses all devices.  Checks that the records of the sharded
// run are bit-identical to the single-device run, in front order, for a batch that does not divide evenly.
// Exit code 0 = ok, 77 = no GPU.  argv[1] = number of fronts (default 37).
#include "../../computing-conjugate-points_b200/host/ccp_frame_det.hpp"
#include "../../include/ccp_setup.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
using namespace ccp_host;
int main(int argc, char** argv) {
    const int n_dev = ccp_device_count();
    if (n_dev <= 0) { std::printf("no GPU: compiled and linked only\n"); return 77; }
    const int M = argc > 1 ? std::atoi(argv[1]) : 37;
    const int Nc = (M + 3) / 4;
    std::vector<double> P((size_t)Nc * 4 * 5); double b3[3] = {1.0, 0.975, 0.95};
    if (ccp_parameter_list(Nc, 4, 0.02, 0.06, b3, P.data()) != 0) return 2;
    std::vector<ccp_front_desc> d(M); std::vector<ccp_setup_info> info(M);
    ccp_setup_opts o{}; o.xy_nbd = -1; o.cone_L = -1; o.eig_err = -1; o.scale = -1; o.L_minus_percent = -1; o.L_minus_override = 3.0;   // short fronts: 96 steps
    if (ccp_setup_sweep(3, P.data(), M, &o, d.data(), info.data(), 0) != 0) return 3;
    // single device
    if (ccp_init(0) != 0) { std::printf("init failed: %s\n", ccp_last_error()); return 1; }
    if (ccp_init(0) != 0) return 4;                                            // same device again: no-op
    if (n_dev > 1 && ccp_init(1) != CCP_E_ARG) { std::printf("ccp_init accepted a second, different device\n"); return 5; }
    std::vector<ccp_front_result> one(M), many(M);
    if (ccp_frame_count_batch(d.data(), M, one.data()) != 0) return 6;
    ccp_shutdown();
    // every device of the box
    if (ccp_init_multi(0) != 0) { std::printf("init_multi failed: %s\n", ccp_last_error()); return 7; }
    if (ccp_bound_devices() != n_dev) return 8;
    if (ccp_frame_count_batch_multi(d.data(), M, many.data()) != 0) { std::printf("multi failed: %s\n", ccp_last_error()); return 9; }
    if (std::memcmp(one.data(), many.data(), sizeof(ccp_front_result) * M) != 0) { std::printf("sharded records differ from the single-device run\n"); return 10; }
    // and through the shim: frameDetBatch uses every bound device
    std::vector<propagateManifold> fronts; std::vector<interval> L;
    for (int i = 0; i < M; i++) { fronts.emplace_back(d[i]); L.push_back(interval(d[i].L_minus)); }
    std::vector<int> codes = propagateManifold::frameDetBatchUnchecked(fronts, L, 14);
    int certified = 0;
    for (int i = 0; i < M; i++) { if (codes[i] != (one[i].failure_count ? -3 : one[i].zero_count)) return 11; certified += codes[i] >= 0; }
    std::printf("multi-GPU ok: %d fronts over %d device(s), records bit-identical to the single-device run, %d certified\n", M, n_dev, certified);
    ccp_shutdown();
    return 0;
}
