// ccp_frame_det.hpp -- C++ host shim above the C ABI (include/ccp.h) with the reference's entry-point names.
//
// The reference's hot path is entered through ONE call,
//     int unstable_e_values = E_u.frameDet(L_minus, grid, endPoint_LPlus - p_s);       heteroclinic.cpp:334
// on a `propagateManifold` built from CAPD objects (propagateManifold.h:63).  This header keeps that shape with
// plain value types (no CAPD): a maintainer replaces the CAPD-typed members by the accessors named below
// (INTEGRATION.md shows the exact lines) and the remaining host pipeline -- stages (i), (ii), (iv) -- keeps working.
//
//   propagateManifold::frameDet      propagateManifold.cpp:125-188   same return codes: >=0 count, -3, -4, -5
//   propagateManifold::plotting      propagateManifold.h:65
//   topFrame::countZeros / getLastFrame / getFirstFrame / makePlot    topFrame.h:75-84, topFrame.cpp:177-351
//   SampleParameters-style batching  heteroclinic.cpp:449-471 -> frameDetBatch (one GPU launch for all fronts)
#pragma once
#include "../../include/ccp.h"
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace ccp_host {

struct interval {
    double lo = 0, hi = 0;
    interval() {}
    interval(double x) : lo(x), hi(x) {}
    interval(double l, double h) : lo(l), hi(h) {}
    double leftBound() const { return lo; }
    double rightBound() const { return hi; }
};
typedef std::vector<interval> IVector;
typedef std::vector<std::vector<interval>> IMatrix;      // row-major
typedef std::vector<std::vector<double>> DMatrix;

// The per-front host algebra of a batched stage (descriptor marshalling, Krawczyk operator, 6x6 interval products, the L_+ test) is independent
// between fronts: with 1,024 fronts it is longer than the launch itself when run on one host thread, so the loops over fronts are
// split over the host cores.  Exceptions must not escape f.
template <class F> inline void parallel_for(size_t n, F f) {
    const size_t hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nt = std::min(hw, n / 16);                 // below 32 fronts a thread costs more than it saves
    if (nt <= 1) { for (size_t i = 0; i < n; i++) f(i); return; }
    std::vector<std::thread> th;
    const size_t chunk = (n + nt - 1) / nt;
    for (size_t t = 0; t < nt; t++) {
        const size_t lo = t * chunk, hi = std::min(n, lo + chunk);
        if (lo >= hi) break;
        th.emplace_back([=] { for (size_t i = lo; i < hi; i++) f(i); });
    }
    for (auto& x : th) x.join();
}

inline ccp_interval to_c(const interval& a) { ccp_interval o; o.lo = a.lo; o.hi = a.hi; return o; }
inline interval from_c(const ccp_interval& a) { return interval(a.lo, a.hi); }

// What topFrame offers after initialize(): built from one result record (+ optional series).
class topFrame {
  public:
    topFrame() {}
    topFrame(int n, const ccp_front_result& r, std::vector<ccp_series_rec> series = {}) : n_(n), res_(r), series_(std::move(series)) {}
    // {zero_count, failure_count}  (topFrame.cpp:193-263)
    std::vector<int> countZeros() const { return {res_.zero_count, res_.failure_count}; }
    // 2n x (n+2): eigen-directions at the right end of the last sub-interval, phi', phi  (topFrame.cpp:265-299)
    IMatrix getLastFrame() const {
        IMatrix M(2 * n_, IVector(n_ + 2));
        for (int i = 0; i < 2 * n_; i++) for (int j = 0; j < n_ + 2; j++) M[i][j] = from_c(res_.last_frame[i * (CCP_MAX_N + 2) + j]);
        return M;
    }
    // 2n x n: all eigen-directions but j_hat at the left end of the first sub-interval, then phi'  (topFrame.cpp:301-331)
    IMatrix getFirstFrame(int j_hat) const {
        IMatrix M(2 * n_, IVector(n_));
        for (int i = 0; i < 2 * n_; i++) {
            int col = 0;
            for (int k = 0; k < n_; k++) { if (k == j_hat) continue; M[i][col++] = from_c(res_.first_frame[i * (CCP_MAX_N + 1) + k]); }
            M[i][n_ - 1] = from_c(res_.first_frame[i * (CCP_MAX_N + 1) + n_]);
        }
        return M;
    }
    // plot_det.txt exactly as topFrame::makePlot + plot() write it (topFrame.cpp:177-191, utils.cpp:126-135)
    void makePlot(const char* path = "plot_det.txt") const {
        FILE* f = std::fopen(path, "w");
        if (!f) throw std::runtime_error("cannot open plot file");
        for (const auto& s : series_)
            std::fprintf(f, "%.16g %.16g %.16g %.16g\n", (s.time.hi + s.time.lo) / 2., (s.det_V.hi + s.det_V.lo) / 2.,
                         (s.time.hi - s.time.lo) / 2., (s.det_V.hi - s.det_V.lo) / 2.);
        std::fclose(f);
    }
    const ccp_front_result& result() const { return res_; }
    const std::vector<ccp_series_rec>& series() const { return series_; }
  private:
    int n_ = 0;
    ccp_front_result res_{};
    std::vector<ccp_series_rec> series_;
};

class propagateManifold {
  public:
    // Host-side stages that stay in the caller's code.  They are part of the proof and therefore MANDATORY: frameDet / frameDetBatch
    // throw std::logic_error while a hook is unset -- a forgotten hook must never turn into a "validated" count.
    //   below_Lminus : localManifold_Eig::checkConjugatePointsBelowLminus   (propagateManifold.cpp:145)      -> -5 when false
    //   L_plus       : propagateManifold::lastEuFrame(A_frame, endPoint_LPlus) (propagateManifold.cpp:170, :192-240) -> -4 when false
    // (host/ccp_local_manifold_eig.hpp and host/ccp_l_plus.hpp provide both without CAPD.)  frameDetUnchecked / frameDetBatchUnchecked
    // are the explicit opt-out for tests and benchmarks of the count alone: they skip both stages and say so in their name.
    std::function<bool()> below_Lminus;
    std::function<bool(const topFrame&, const IVector&)> L_plus;

    // The data construct_InitCondU / construct_A_lin read from the CAPD objects (propagateManifold.cpp:6-81):
    //   params            All_parameters_interval          (b_1..b_n, c_12..)
    //   A_u               (*(*pUnstable).pF).A             point matrix 2n x 2n
    //   p_i, Uxy          (*pUnstable).getPointNbd(XY_pt, XY_nbd)
    //   Eu_m_Error_Final  after (*pUnstable).computeEigenError_minus_infty()
    propagateManifold(int n, const IVector& params, const DMatrix& A_u, const IVector& p_i, const IVector& Uxy,
                      const IMatrix& Eu_m_Error_Final, int order, int step_size) {
        if (n < 2 || n > CCP_MAX_N) throw std::invalid_argument("n out of range");
        desc_ = ccp_front_desc{};
        desc_.n = n; desc_.order = order; desc_.step_log2 = step_size;
        for (int i = 0; i < n; i++) desc_.b[i] = to_c(params[i]);
        for (int i = 0; i < n - 1; i++) desc_.c[i] = to_c(params[n + i]);
        for (int i = 0; i < 2 * n; i++) {
            desc_.p_i[i] = to_c(p_i[i]); desc_.Uxy[i] = to_c(Uxy[i]);
            for (int j = 0; j < 2 * n; j++) desc_.A_u[i * CCP_MAX_DIM + j] = A_u[i][j];
        }
        for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) desc_.Eu_err[j * CCP_MAX_N + k] = to_c(Eu_m_Error_Final[j][k]);
    }
    explicit propagateManifold(const ccp_front_desc& d) : desc_(d) {}

    void plotting(bool make_plot_input) { MAKE_PLOT = make_plot_input; }

    // Same contract as the reference: count >= 0, -3 count failed, -4 L_+ failed, -5 below-L_- failed;
    // an engine / enclosure failure throws (the reference's CAPD would throw too; SampleParameters maps it to -10).
    int frameDet(interval L_minus, int grid, const IVector& endPoint_LPlus) {
        require_hooks();
        if (!below_Lminus()) return -5;
        desc_.L_minus = L_minus.hi; desc_.grid = grid;
        ccp_front_result r{};
        std::vector<ccp_series_rec> series;
        if (MAKE_PLOT) {
            size_t len = 0;
            check(ccp_frame_count_series(&desc_, &r, nullptr, 0, &len));
            series.resize(len);
            check(ccp_frame_count_series(&desc_, &r, series.data(), series.size(), &len));
        } else check(ccp_frame_count_batch(&desc_, 1, &r));
        return finish(r, std::move(series), endPoint_LPlus);
    }
    // The frame count WITHOUT the two host-side proof stages (Prop. 2.5 below L_-, the L_+ condition): not a validated result.
    int frameDetUnchecked(interval L_minus, int grid) {
        below_Lminus = [] { return true; };
        L_plus = [](const topFrame&, const IVector&) { return true; };
        return frameDet(L_minus, grid, IVector());
    }

    // The sweep of SampleParameters in one launch: all fronts go to the GPU together.
    // `parallel_hooks`: the below-L_- and L_+ hooks of different fronts may run concurrently (they must then only touch per-front state;
    // the L_+ test is an interval eigen-computation per front and dominates this call for a large sweep)
    static std::vector<int> frameDetBatch(std::vector<propagateManifold>& fronts, const std::vector<interval>& L_minus, int grid,
                                          const std::vector<IVector>& endPoints, bool parallel_hooks = false) {
        std::vector<ccp_front_desc> d(fronts.size());
        std::vector<ccp_front_result> r(fronts.size());
        std::vector<int> out(fronts.size(), -10);
        for (auto& f : fronts) f.require_hooks();
        for (size_t i = 0; i < fronts.size(); i++) { fronts[i].desc_.L_minus = L_minus[i].hi; fronts[i].desc_.grid = grid; d[i] = fronts[i].desc_; }
        // every GPU bound by ccp_init_multi takes part (front f on device f mod n_devices); with ccp_init it is one device
        check(ccp_frame_count_batch_multi(d.data(), d.size(), r.data()));
        auto one = [&](size_t i) {
            if (!fronts[i].below_Lminus()) { out[i] = -5; return; }
            try { out[i] = fronts[i].finish(r[i], {}, endPoints[i]); } catch (const std::exception&) { out[i] = -10; }
        };
        if (parallel_hooks) parallel_for(fronts.size(), one); else for (size_t i = 0; i < fronts.size(); i++) one(i);
        return out;
    }
    static std::vector<int> frameDetBatchUnchecked(std::vector<propagateManifold>& fronts, const std::vector<interval>& L_minus, int grid) {
        for (auto& f : fronts) { f.below_Lminus = [] { return true; }; f.L_plus = [](const topFrame&, const IVector&) { return true; }; }
        return frameDetBatch(fronts, L_minus, grid, std::vector<IVector>(fronts.size()));
    }
    const topFrame& frame() const { return frame_; }
    const ccp_front_desc& descriptor() const { return desc_; }

  private:
    int finish(const ccp_front_result& r, std::vector<ccp_series_rec> series, const IVector& endPoint_LPlus) {
        if (r.status != CCP_OK) throw std::runtime_error("ccp: front status " + std::to_string(r.status));
        frame_ = topFrame(desc_.n, r, std::move(series));
        if (MAKE_PLOT) frame_.makePlot();
        if (!L_plus(frame_, endPoint_LPlus)) return -4;
        std::vector<int> cp = frame_.countZeros();
        return cp[1] > 0 ? -3 : cp[0];
    }
    void require_hooks() const {
        if (!below_Lminus || !L_plus)
            throw std::logic_error("propagateManifold: the below-L_- and L_+ stages are not wired (set below_Lminus and L_plus, or call frameDetUnchecked)");
    }
    static void check(int rc) { if (rc != 0) throw std::runtime_error(std::string("ccp: ") + ccp_last_error()); }
    ccp_front_desc desc_{};
    topFrame frame_;
    bool MAKE_PLOT = false;
};

} // namespace ccp_host
