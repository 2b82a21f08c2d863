// computeFrontAndConjugatePoints (heteroclinic.cpp:23-347) through the stage mirrors, for the parameters of the reference's main()
// (n = 3: b = 1, .98, .96, c = .04, .02; heteroclinic.cpp:513-531): expected result = 2 conjugate points.  0 = ok, 77 = no GPU.
#include "../../computing-conjugate-points_b200/host/ccp_driver.hpp"
#include <cstdio>
#include <cstdlib>
using namespace ccp_host;
// usage: driver_gpu [expected n p_1 ... p_{2n-1}]      default: 2 3 1 .98 .96 .04 .02
int main(int argc, char** argv) {
    if (ccp_device_count() <= 0) { std::printf("no GPU: compiled and linked only\n"); return 77; }
    if (ccp_init(0) != 0) { std::printf("init failed: %s\n", ccp_last_error()); return 2; }
    std::vector<double> par = {1.0, 0.98, 0.96, 0.04, 0.02};
    int expected = 2, dim = 6;
    if (argc > 2) {
        expected = std::atoi(argv[1]); dim = 2 * std::atoi(argv[2]);
        if (argc != 3 + dim - 1) { std::printf("expected %d parameters\n", dim - 1); return 3; }
        par.clear();
        for (int i = 3; i < argc; i++) par.push_back(std::atof(argv[i]));
    }
    IVector ipar(par.size());
    for (size_t i = 0; i < par.size(); i++) ipar[i] = interval(std::nextafter(par[i], -INFINITY), std::nextafter(par[i], INFINITY));   // interval("x","x")
    DriverReport rep;
    int r = -10;
    try { r = computeFrontAndConjugatePoints(dim, par, ipar, &rep); }
    catch (const std::exception& e) { std::printf("exception: %s\n", e.what()); }
    catch (int code) { std::printf("thrown code %d\n", code); r = code; }
    std::printf("manifolds %d/%d, T = %.6f, inside %d, proof %d (radius %.3e), eps_- = %.3e, below L_- %d, L_- = %.6f\n", (int)rep.conditions_first,
                (int)rep.conditions_resized, rep.T, (int)rep.inside_box, (int)rep.proof, rep.radius, rep.eps_minus, (int)rep.below_Lminus, rep.L_minus.hi);
    std::printf("countZeros: %d zeros, %d failures (first at sub-interval %d of %d)\n", rep.zero_count, rep.failure_count, rep.first_failure, rep.series_length);
    std::printf("computeFrontAndConjugatePoints = %d\n", r);
    return r == expected ? 0 : 1;
}
