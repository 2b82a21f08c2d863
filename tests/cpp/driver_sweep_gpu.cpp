// SampleParameters (heteroclinic.cpp:429-491) with the complete proof per front, all fronts in lock-step launches:
// computeFrontAndConjugatePointsBatch over Construct_ParameterList(N_c, N_r); a few fronts are repeated with the single-front
// driver and must give the same code.  usage: driver_sweep_gpu [N_c N_r]   (default 8 2).  0 = ok, 77 = no GPU.
#include "../../computing-conjugate-points_b200/host/ccp_driver.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
using namespace ccp_host;
int main(int argc, char** argv) {
    if (ccp_device_count() <= 0) { std::printf("no GPU: compiled and linked only\n"); return 77; }
    if (ccp_init(0) != 0) { std::printf("init failed: %s\n", ccp_last_error()); return 2; }
    if (argc == 2 && std::string(argv[1]) == "n2") {                             // three fronts of the two-component system (dimension 4)
        const std::vector<std::vector<double>> P = {{1, .97, .05}, {1, .97, -.05}, {1, .95, .03}};
        std::vector<IVector> PI(P.size());
        for (size_t q = 0; q < P.size(); q++) for (double v : P[q]) PI[q].push_back(interval(v));
        std::vector<DriverReport> rep;
        const std::vector<int> code = computeFrontAndConjugatePointsBatch(4, P, PI, &rep);
        std::printf("n = 2 codes:");
        for (size_t q = 0; q < code.size(); q++) std::printf(" %d", code[q]);
        std::printf("\n");
        for (size_t q = 0; q < code.size(); q++) {
            if (!rep[q].proof || rep[q].failure_count != 0 || rep[q].zero_count != code[q]) return 5;
            if (computeFrontAndConjugatePoints(4, P[q], PI[q]) != code[q]) return 6;
        }
        return (code[0] == 1 && code[1] == 0 && code[2] == 1) ? 0 : 7;
    }
    const int N_c = argc > 2 ? std::atoi(argv[1]) : 8, N_r = argc > 2 ? std::atoi(argv[2]) : 2, m = N_c * N_r;
    const double b[3] = {1.0, 0.975, 0.95};
    std::vector<double> flat(5 * m);
    if (ccp_parameter_list(N_c, N_r, 0.02, 0.06, b, flat.data()) != 0) return 3;
    std::vector<std::vector<double>> P(m); std::vector<IVector> PI(m);
    for (int q = 0; q < m; q++) { P[q].assign(flat.begin() + 5 * q, flat.begin() + 5 * q + 5); for (double v : P[q]) PI[q].push_back(interval(v)); }
    std::vector<DriverReport> rep;
    const auto t0 = std::chrono::steady_clock::now();
    const std::vector<int> code = computeFrontAndConjugatePointsBatch(6, P, PI, &rep);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int hist[8] = {0}, proven = 0;
    for (int q = 0; q < m; q++) { if (code[q] >= 0 && code[q] < 8) hist[code[q]]++; proven += rep[q].proof ? 1 : 0; }
    std::printf("sweep %d x %d: %d fronts, %d connecting orbits proven, counts 0/1/2: %d/%d/%d, failed: %d, %.2f s\n", N_c, N_r, m, proven, hist[0], hist[1], hist[2],
                m - hist[0] - hist[1] - hist[2], sec);
    for (const auto& st : driverStageTimes()) std::printf("  stage %-44s %8.3f s\n", st.first.c_str(), st.second);
    for (int q = 0; q < m; q++) std::printf("%s%d", q ? " " : "codes: ", code[q]);
    std::printf("\n");
    int bad = 0;
    for (int q = 0; q < m; q += std::max(1, m / 3)) {
        int r = -10;
        try { r = computeFrontAndConjugatePoints(6, P[q], PI[q]); } catch (const std::exception&) { r = -10; } catch (int e) { r = e; }
        if (r != code[q]) { std::printf("front %d: single-front driver %d, batch %d\n", q, r, code[q]); bad++; }
    }
    if (bad) return 1;
    std::printf("batch == single-front driver on the probed fronts\n");
    return proven > 0 ? 0 : 4;                        // some fronts legitimately fail a stage with the default constants (codes above)
}
