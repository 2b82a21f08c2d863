// Writes plot_det.txt through the C++ shim exactly as the reference's driver does (heteroclinic.cpp:61, 299-334: plotting(true), then
// frameDet -> topFrame::makePlot, topFrame.cpp:177-191) for the parameters and L_- of the reference's recorded run.
// Exit code 0 = ok, 77 = no GPU.  The file lands in the current directory, like the reference's.
#include "../../computing-conjugate-points_b200/host/ccp_frame_det.hpp"
#include "../../include/ccp_setup.h"
#include <cmath>
#include <cstdio>
using namespace ccp_host;
int main() {
    if (ccp_device_count() <= 0) { std::printf("no GPU\n"); return 77; }
    if (ccp_init(0) != 0) { std::printf("init failed: %s\n", ccp_last_error()); return 1; }
    double params[5] = {1.0, 0.98, 0.96, 0.04, 0.02};                      // heteroclinic.cpp:524-529
    ccp_front_desc d; ccp_setup_info info;
    ccp_setup_opts o{}; o.xy_nbd = -1; o.cone_L = -1; o.eig_err = -1; o.scale = -1; o.L_minus_percent = -1;
    o.L_minus_override = 33.84706312634513;                                // last t_mid + t_rad of the reference's plot_det.txt
    if (ccp_setup_front(3, params, &o, &d, &info) != 0) return 2;
    for (int i = 0; i < 3; i++) { d.b[i].lo = std::nextafter(d.b[i].lo, -INFINITY); d.b[i].hi = std::nextafter(d.b[i].hi, INFINITY); }   // interval("x","x")
    for (int i = 0; i < 2; i++) { d.c[i].lo = std::nextafter(d.c[i].lo, -INFINITY); d.c[i].hi = std::nextafter(d.c[i].hi, INFINITY); }
    propagateManifold E_u(d);
    E_u.plotting(true);
    const int count = E_u.frameDetUnchecked(interval(d.L_minus), 14);
    std::printf("frameDet = %d, %zu rows written to plot_det.txt\n", count, E_u.frame().series().size());
    return count == 2 ? 0 : 3;
}
