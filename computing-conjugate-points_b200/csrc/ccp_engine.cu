// ccp_engine.cu -- C ABI of libccp_b200.so (include/ccp.h): engine life cycle, batched launches,
// FP64 peak microbenchmark.  There is NO CPU fallback: without a CUDA device every compute entry
// point returns CCP_E_NO_DEVICE.
#include "ccp_kernel.cuh"
#include "ccp_dfu.cuh"
#include "ccp_subdiv.cuh"
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cstddef>
#include <mutex>
#include <string>
#include <cstdlib>

namespace {

// One context per bound GPU.  ccp_init binds ONE device (context 0); ccp_init_multi binds devices 0..n-1 and
// ccp_frame_count_batch_multi shards a batch over them (front_id mod n_devices) from this one host thread.
struct Dev {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void* d_desc = nullptr;  size_t desc_cap = 0;
    void* d_res = nullptr;   size_t res_cap = 0;
    int*  d_index = nullptr; size_t index_cap = 0;
    void* d_series = nullptr; size_t series_cap = 0;
    void* h_desc = nullptr;  size_t h_desc_cap = 0;      // pinned staging of this device's shard (multi-GPU entry point)
    void* h_res = nullptr;   size_t h_res_cap = 0;
    void* d_sub = nullptr;   size_t sub_cap = 0;         // config 4: sub-box descriptors / records / per-parameter hulls
    void* d_subres = nullptr; size_t subres_cap = 0;
    void* d_hull = nullptr;  size_t hull_cap = 0;
    int grid_limit[CCP_MAX_N + 1] = {0, 0, 0, 0, 0};
};
struct Engine {
    bool ready = false;
    std::vector<Dev> devs;
    unsigned long long launches = 0;
    float last_ms = 0.f;
    std::string err;
    std::mutex mu;
};
Engine E;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { E.err = std::string(#call) + ": " + cudaGetErrorString(e_); return CCP_E_CUDA; } } while (0)

// descriptors whose step count would not fit the kernel's 32-bit counters are rejected before launch (the kernel answers
// CCP_ST_BAD_DESC for them as well): L_minus * 2^step_log2 <= 2^24 steps
constexpr double MAX_STEPS = 16777216.0;

template <int N> int prepare_kernel(Dev& dv) {
    const size_t smem = sizeof(ccp::FrontSmem<N>);
    CK(cudaFuncSetAttribute(ccp::front_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ccp::front_kernel<N>, ccp::NT, smem));
    if (per_sm < 1) { E.err = "kernel does not fit on an SM"; return CCP_E_CUDA; }
    dv.grid_limit[N] = per_sm * dv.sm_count;
    return 0;
}

template <int N> int launch(Dev& dv, const ccp_front_desc* d_desc, ccp_front_result* d_res, const int* d_index, int count,
                            ccp_series_rec* series, int series_cap, cudaStream_t st) {
    if (count <= 0) return 0;
    // one CTA of ccp::NT threads per front; the grid is sized to the number of CTAs that are resident at once
    // (multiple of the SM count) and strides over the batch.
    int grid = count < dv.grid_limit[N] ? count : dv.grid_limit[N];
    static const int dbg_align = getenv("CCP_DBG_ALIGN") ? atoi(getenv("CCP_DBG_ALIGN")) : 17;         // experiment knobs (scripts/): bit 0 = per-SM alignment, bit 4 = warp roles exchanged in every second CTA of a sub-partition pair
    static const int dbg_pad = getenv("CCP_DBG_SMEM_PAD") ? atoi(getenv("CCP_DBG_SMEM_PAD")) : 0;
    if (dbg_pad) { cudaFuncSetAttribute(ccp::front_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ccp::FrontSmem<N>) + dbg_pad); }
    ccp::front_kernel<N><<<grid, ccp::NT, sizeof(ccp::FrontSmem<N>) + dbg_pad, st>>>(d_desc, d_res, d_index, count, series, series_cap, dbg_align);
    E.launches++;
    CK(cudaGetLastError());
    return 0;
}

int ensure(void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    CK(cudaMalloc(p, bytes));
    *cap = bytes;
    return 0;
}
int ensure_pinned(void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return 0;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    CK(cudaMallocHost(p, bytes));
    *cap = bytes;
    return 0;
}

int launch_n(Dev& dv, int n, const ccp_front_desc* d_desc, ccp_front_result* d_res, const int* d_index, int count, cudaStream_t st) {
    switch (n) {
        case 2: return launch<2>(dv, d_desc, d_res, d_index, count, nullptr, 0, st);
        case 3: return launch<3>(dv, d_desc, d_res, d_index, count, nullptr, 0, st);
        case 4: return launch<4>(dv, d_desc, d_res, d_index, count, nullptr, 0, st);
        default: E.err = "unsupported n"; return CCP_E_ARG;
    }
}

// launches the per-n kernels for a device-resident batch; `h_n` = host copy of the n field of every descriptor
int run_device(Dev& dv, const ccp_front_desc* d_desc, ccp_front_result* d_res, const int* h_n, size_t nf, cudaStream_t st) {
    bool uniform = true;
    for (size_t i = 1; i < nf; i++) if (h_n[i] != h_n[0]) { uniform = false; break; }
    if (uniform && h_n[0] >= 2 && h_n[0] <= 4) return launch_n(dv, h_n[0], d_desc, d_res, nullptr, (int)nf, st);
    // mixed batch: one index list per n; anything else is answered by the n=2 kernel with CCP_ST_BAD_DESC
    std::vector<int> idx[3];
    for (size_t i = 0; i < nf; i++) { int n = h_n[i]; idx[(n == 3) ? 1 : (n == 4) ? 2 : 0].push_back((int)i); }
    int rc = ensure((void**)&dv.d_index, &dv.index_cap, nf * sizeof(int)); if (rc) return rc;
    size_t off = 0;
    for (int k = 0; k < 3; k++) {
        if (idx[k].empty()) continue;
        CK(cudaMemcpyAsync(dv.d_index + off, idx[k].data(), idx[k].size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));   // idx[k] is a stack-owned vector
        if ((rc = launch_n(dv, k == 0 ? 2 : k == 1 ? 3 : 4, d_desc, d_res, dv.d_index + off, (int)idx[k].size(), st))) return rc;
        off += idx[k].size();
    }
    return 0;
}

int init_device(Dev& dv, int device) {
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    dv.device = device; dv.sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&dv.stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&dv.ev0));
    CK(cudaEventCreate(&dv.ev1));
    int rc;
    if ((rc = prepare_kernel<2>(dv))) return rc;
    if ((rc = prepare_kernel<3>(dv))) return rc;
    if ((rc = prepare_kernel<4>(dv))) return rc;
    {   // reciprocal tables of the front kernel (constant memory of THIS device), computed by the device's own directed divisions
        constexpr int NI = ccp::KMAX + 2;
        double* d_tab = nullptr;
        CK(cudaMalloc(&d_tab, sizeof(double) * 3 * NI));
        ccp::inv_table_kernel<<<1, 64, 0, dv.stream>>>(d_tab);
        CK(cudaGetLastError());
        CK(cudaMemcpyToSymbolAsync(ccp::c_invlo, d_tab, sizeof(double) * NI, 0, cudaMemcpyDeviceToDevice, dv.stream));
        CK(cudaMemcpyToSymbolAsync(ccp::c_invhi, d_tab + NI, sizeof(double) * NI, 0, cudaMemcpyDeviceToDevice, dv.stream));
        CK(cudaMemcpyToSymbolAsync(ccp::c_invs, d_tab + 2 * NI, sizeof(double) * NI, 0, cudaMemcpyDeviceToDevice, dv.stream));
        CK(cudaStreamSynchronize(dv.stream));
        CK(cudaFree(d_tab));
    }
    // per-lane constants of the table loop (ccp_kernel.cuh: g_lanec), one table per n
    ccp::lanec_table_kernel<2><<<1, ccp::NT, 0, dv.stream>>>();
    ccp::lanec_table_kernel<3><<<1, ccp::NT, 0, dv.stream>>>();
    ccp::lanec_table_kernel<4><<<1, ccp::NT, 0, dv.stream>>>();
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(dv.stream));
    CK(cudaFuncSetAttribute(ccp::dfu_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ccp::DfuSmem<2>)));
    CK(cudaFuncSetAttribute(ccp::dfu_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ccp::DfuSmem<3>)));
    CK(cudaFuncSetAttribute(ccp::dfu_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ccp::DfuSmem<4>)));
    return 0;
}
void free_device(Dev& dv) {
    if (dv.device < 0) return;
    cudaSetDevice(dv.device);
    if (dv.stream) cudaStreamSynchronize(dv.stream);
    if (dv.d_desc) cudaFree(dv.d_desc);
    if (dv.d_res) cudaFree(dv.d_res);
    if (dv.d_index) cudaFree(dv.d_index);
    if (dv.d_series) cudaFree(dv.d_series);
    if (dv.d_sub) cudaFree(dv.d_sub);
    if (dv.d_subres) cudaFree(dv.d_subres);
    if (dv.d_hull) cudaFree(dv.d_hull);
    if (dv.h_desc) cudaFreeHost(dv.h_desc);
    if (dv.h_res) cudaFreeHost(dv.h_res);
    if (dv.ev0) cudaEventDestroy(dv.ev0);
    if (dv.ev1) cudaEventDestroy(dv.ev1);
    if (dv.stream) cudaStreamDestroy(dv.stream);
    dv = Dev();
}
// the single-device entry points run on context 0
#define DEV0() Dev& dv = E.devs[0]; CK(cudaSetDevice(dv.device))

// ---- FP64 peak: independent DFMA chains, 8 per thread ----
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            x0 = __fma_ru(x0, a, b); x1 = __fma_rd(x1, a, b); x2 = __fma_ru(x2, a, b); x3 = __fma_rd(x3, a, b);
            x4 = __fma_ru(x4, a, b); x5 = __fma_rd(x5, a, b); x6 = __fma_ru(x6, a, b); x7 = __fma_rd(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// ---- primitive probe: one op per element, 8 inputs -> 4 outputs (unit tests compare with the oracle bit for bit) ----
__global__ void debug_prim_kernel(int op, const double* in, double* out, int n) {
    using namespace ccp;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* x = in + 8 * (size_t)i; double* o = out + 4 * (size_t)i;
    o[0] = o[1] = o[2] = o[3] = 0.0;
    switch (op) {
        case 0: o[0] = __dadd_rd(x[0], x[1]); break;
        case 1: o[0] = __dadd_ru(x[0], x[1]); break;
        case 2: o[0] = __dmul_rd(x[0], x[1]); break;
        case 3: o[0] = __dmul_ru(x[0], x[1]); break;
        case 4: o[0] = __fma_rd(x[0], x[1], x[2]); break;
        case 5: o[0] = __fma_ru(x[0], x[1], x[2]); break;
        case 6: o[0] = __ddiv_rd(x[0], x[1]); break;
        case 7: o[0] = __ddiv_ru(x[0], x[1]); break;
        case 10: { ival r = imul(mk(x[0], x[1]), mk(x[2], x[3])); o[0] = r.lo; o[1] = r.hi; break; }
        case 11: { mt r = iv2mt(mk(x[0], x[1])); o[0] = r.m; o[1] = r.t; break; }
        case 12: { mt a; a.m = x[0]; a.t = x[1]; ival r = mt2iv(a); o[0] = r.lo; o[1] = r.hi; break; }
        case 13: { acc4 a = acc_zero(); mt p, q; p.m = x[0]; p.t = x[1]; q.m = x[2]; q.t = x[3]; acc_term(a, p, q);
                   p.m = x[4]; p.t = x[5]; q.m = x[6]; q.t = x[7]; acc_term_neg(a, p, q); ival r = acc_finish(a); o[0] = r.lo; o[1] = r.hi; break; }
        case 14: { double il[32], ih[32]; const int k = (int)x[2]; il[k] = __ddiv_rd(1.0, (double)k); ih[k] = __ddiv_ru(1.0, (double)k);
                   ival r = idivk_tab(mk(x[0], x[1]), il, ih, k); o[0] = r.lo; o[1] = r.hi; break; }
        case 15: { ival r = horner_rg(mk(x[0], x[1]), x[2], x[3], mk(x[4], x[5])); o[0] = r.lo; o[1] = r.hi; break; }
        case 16: { ival r = isub(mk(x[0], x[1]), mk(x[2], x[3])); o[0] = r.lo; o[1] = r.hi; break; }
        case 17: { ival r = iadd(mk(x[0], x[1]), mk(x[2], x[3])); o[0] = r.lo; o[1] = r.hi; break; }
        case 18: { ival r = imulk(mk(x[0], x[1]), (int)x[2]); o[0] = r.lo; o[1] = r.hi; break; }
        case 19: { ival r = horner_pt(mk(x[0], x[1]), x[2], mk(x[4], x[5])); o[0] = r.lo; o[1] = r.hi; break; }
        case 20: { ival m[3][3]; for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { int q = (a * 3 + b) % 7; m[a][b] = mk(dmin(x[q], x[q + 1]), dmax(x[q], x[q + 1])); }
                   ival r = Det<3>::det(m); o[0] = r.lo; o[1] = r.hi; break; }
        case 21: { const int k = (int)x[2]; const double ih = __ddiv_ru(1.0, (double)k);
                   mt a; a.m = x[0]; a.t = x[1]; mt r = mt_scale(a, ih, __dmul_ru(ih, 1.0000000000000009)); o[0] = r.m; o[1] = r.t; break; }
        case 22: { ival a = mk(x[0], x[1]); ipfma(a, mk(x[2], x[3]), x[4]); o[0] = a.lo; o[1] = a.hi; break; }
        default: break;
    }
}

} // namespace

extern "C" {

int ccp_debug_primitives(int op, const double* in, double* out, size_t n) {
    if (!E.ready) { E.err = "ccp_init not called (or no device)"; return CCP_E_NOT_INIT; }
    std::lock_guard<std::mutex> lk(E.mu);
    DEV0();
    double *d_in = nullptr, *d_out = nullptr;
    CK(cudaMalloc(&d_in, n * 8 * sizeof(double)));
    CK(cudaMalloc(&d_out, n * 4 * sizeof(double)));
    CK(cudaMemcpyAsync(d_in, in, n * 8 * sizeof(double), cudaMemcpyHostToDevice, dv.stream));
    debug_prim_kernel<<<(unsigned)((n + 127) / 128), 128, 0, dv.stream>>>(op, d_in, d_out, (int)n);
    E.launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_out, n * 4 * sizeof(double), cudaMemcpyDeviceToHost, dv.stream));
    CK(cudaStreamSynchronize(dv.stream));
    cudaFree(d_in); cudaFree(d_out);
    return 0;
}

#ifdef CCP_TRACE
int ccp_debug_trace(long long* out, int* n) {
    cudaMemcpyFromSymbol(out, ccp::g_trace, sizeof(long long) * 1024);
    cudaMemcpyFromSymbol(n, ccp::g_trace_n, sizeof(int) * 2);
    return 0;
}
#endif
#ifdef CCP_PROFILE
int ccp_debug_profile(unsigned long long* out, int reset) {
    if (out) cudaMemcpyFromSymbol(out, ccp::g_prof, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(ccp::g_prof, z, sizeof(z)); }
    return 0;
}
#endif

const char* ccp_version(void) { return "ccp_b200 0.1 (sm_100a, c1-structured interval Taylor kernel)"; }
const char* ccp_last_error(void) { return E.err.c_str(); }

int ccp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ccp_init(int device) {
    std::lock_guard<std::mutex> lk(E.mu);
    if (E.ready) {
        if (E.devs.size() == 1 && E.devs[0].device == device) return 0;
        E.err = "ccp_init: the engine is already bound (device " + std::to_string(E.devs[0].device) + (E.devs.size() > 1 ? ", multi-GPU" : "") +
                "); call ccp_shutdown before binding another device";
        return CCP_E_ARG;
    }
    int n = ccp_device_count();
    if (n <= 0) { E.err = "no CUDA device visible: libccp_b200 has no CPU fallback"; return CCP_E_NO_DEVICE; }
    if (device < 0 || device >= n) { E.err = "bad device index"; return CCP_E_ARG; }
    E.devs.assign(1, Dev());
    int rc = init_device(E.devs[0], device);
    if (rc) { free_device(E.devs[0]); E.devs.clear(); return rc; }
    E.ready = true;
    return 0;
}

int ccp_init_multi(int n_devices) {
    std::lock_guard<std::mutex> lk(E.mu);
    int n = ccp_device_count();
    if (n <= 0) { E.err = "no CUDA device visible: libccp_b200 has no CPU fallback"; return CCP_E_NO_DEVICE; }
    if (n_devices <= 0) n_devices = n;
    if (n_devices > n) { E.err = "ccp_init_multi: more devices requested than visible"; return CCP_E_ARG; }
    if (E.ready) {
        bool same = (int)E.devs.size() == n_devices;
        for (int d = 0; same && d < n_devices; d++) same = E.devs[d].device == d;
        if (same) return 0;
        E.err = "ccp_init_multi: the engine is already bound to another device set; call ccp_shutdown first";
        return CCP_E_ARG;
    }
    E.devs.assign(n_devices, Dev());
    for (int d = 0; d < n_devices; d++) {
        int rc = init_device(E.devs[d], d);
        if (rc) { for (auto& dv : E.devs) free_device(dv); E.devs.clear(); return rc; }
    }
    cudaSetDevice(E.devs[0].device);
    E.ready = true;
    return 0;
}

int ccp_bound_devices(void) { return E.ready ? (int)E.devs.size() : 0; }

void ccp_shutdown(void) {
    std::lock_guard<std::mutex> lk(E.mu);
    if (!E.ready) return;
    for (auto& dv : E.devs) free_device(dv);
    E.devs.clear();
    E.ready = false;
}

int ccp_frame_count_batch(const ccp_front_desc* fronts, size_t n_fronts, ccp_front_result* results) {
    if (!E.ready) { E.err = "ccp_init not called (or no device)"; return CCP_E_NOT_INIT; }
    if (n_fronts == 0) return 0;
    if (!fronts || !results) return CCP_E_ARG;
    std::lock_guard<std::mutex> lk(E.mu);
    DEV0();
    int rc;
    if ((rc = ensure(&dv.d_desc, &dv.desc_cap, n_fronts * sizeof(ccp_front_desc)))) return rc;
    if ((rc = ensure(&dv.d_res, &dv.res_cap, n_fronts * sizeof(ccp_front_result)))) return rc;
    std::vector<int> h_n(n_fronts);
    for (size_t i = 0; i < n_fronts; i++) h_n[i] = fronts[i].n;
    CK(cudaMemcpyAsync(dv.d_desc, fronts, n_fronts * sizeof(ccp_front_desc), cudaMemcpyHostToDevice, dv.stream));
    CK(cudaEventRecord(dv.ev0, dv.stream));
    if ((rc = run_device(dv, (const ccp_front_desc*)dv.d_desc, (ccp_front_result*)dv.d_res, h_n.data(), n_fronts, dv.stream))) return rc;
    CK(cudaEventRecord(dv.ev1, dv.stream));
    CK(cudaMemcpyAsync(results, dv.d_res, n_fronts * sizeof(ccp_front_result), cudaMemcpyDeviceToHost, dv.stream));
    CK(cudaStreamSynchronize(dv.stream));
    CK(cudaEventElapsedTime(&E.last_ms, dv.ev0, dv.ev1));
    return 0;
}

// The sweep of SampleParameters (heteroclinic.cpp:449-471) over every bound GPU from ONE host thread: front f goes to device
// f mod n_devices (interleaved: balances the +-4 % spread of step counts along the parameter list), one stream per device, no
// exchange between devices; the result records land in the caller's array in front order.
int ccp_frame_count_batch_multi(const ccp_front_desc* fronts, size_t n_fronts, ccp_front_result* results) {
    if (!E.ready) { E.err = "ccp_init not called (or no device)"; return CCP_E_NOT_INIT; }
    if (E.devs.size() == 1) return ccp_frame_count_batch(fronts, n_fronts, results);
    if (n_fronts == 0) return 0;
    if (!fronts || !results) return CCP_E_ARG;
    std::lock_guard<std::mutex> lk(E.mu);
    const size_t nd = E.devs.size();
    std::vector<std::vector<int>> h_n(nd);
    int rc;
    for (size_t d = 0; d < nd; d++) {
        Dev& dv = E.devs[d];
        const size_t m = (n_fronts + nd - 1 - d) / nd;                 // fronts d, d + nd, ...
        if (m == 0) continue;
        CK(cudaSetDevice(dv.device));
        if ((rc = ensure(&dv.d_desc, &dv.desc_cap, m * sizeof(ccp_front_desc)))) return rc;
        if ((rc = ensure(&dv.d_res, &dv.res_cap, m * sizeof(ccp_front_result)))) return rc;
        if ((rc = ensure_pinned(&dv.h_desc, &dv.h_desc_cap, m * sizeof(ccp_front_desc)))) return rc;
        if ((rc = ensure_pinned(&dv.h_res, &dv.h_res_cap, m * sizeof(ccp_front_result)))) return rc;
        ccp_front_desc* hd = (ccp_front_desc*)dv.h_desc;
        h_n[d].resize(m);
        for (size_t k = 0; k < m; k++) { hd[k] = fronts[d + k * nd]; h_n[d][k] = hd[k].n; }
        CK(cudaMemcpyAsync(dv.d_desc, hd, m * sizeof(ccp_front_desc), cudaMemcpyHostToDevice, dv.stream));
        CK(cudaEventRecord(dv.ev0, dv.stream));
        if ((rc = run_device(dv, (const ccp_front_desc*)dv.d_desc, (ccp_front_result*)dv.d_res, h_n[d].data(), m, dv.stream))) return rc;
        CK(cudaEventRecord(dv.ev1, dv.stream));
        CK(cudaMemcpyAsync(dv.h_res, dv.d_res, m * sizeof(ccp_front_result), cudaMemcpyDeviceToHost, dv.stream));
    }
    float worst = 0.f;
    for (size_t d = 0; d < nd; d++) {
        Dev& dv = E.devs[d];
        const size_t m = (n_fronts + nd - 1 - d) / nd;
        if (m == 0) continue;
        CK(cudaSetDevice(dv.device));
        CK(cudaStreamSynchronize(dv.stream));
        float ms = 0.f; CK(cudaEventElapsedTime(&ms, dv.ev0, dv.ev1));
        if (ms > worst) worst = ms;
        const ccp_front_result* hr = (const ccp_front_result*)dv.h_res;
        for (size_t k = 0; k < m; k++) results[d + k * nd] = hr[k];
    }
    E.last_ms = worst;                                                  // slowest device
    CK(cudaSetDevice(E.devs[0].device));
    return 0;
}

int ccp_frame_count_batch_device_n(const void* d_fronts, size_t n_fronts, void* d_results, void* stream, int n) {
    if (!E.ready) { E.err = "ccp_init not called (or no device)"; return CCP_E_NOT_INIT; }
    if (n_fronts == 0) return 0;
    if (!d_fronts || !d_results) return CCP_E_ARG;
    std::lock_guard<std::mutex> lk(E.mu);
    DEV0();
    return launch_n(dv, n, (const ccp_front_desc*)d_fronts, (ccp_front_result*)d_results, nullptr, (int)n_fronts, stream ? (cudaStream_t)stream : dv.stream);
}

int ccp_frame_count_batch_device(const void* d_fronts, size_t n_fronts, void* d_results, void* stream) {
    if (!E.ready) { E.err = "ccp_init not called (or no device)"; return CCP_E_NOT_INIT; }
    if (n_fronts == 0) return 0;
    if (!d_fronts || !d_results) return CCP_E_ARG;
    // the n field is needed on the host to pick the kernel instantiation: read the first descriptor's header (one 4-byte copy and
    // a stream synchronisation; ccp_frame_count_batch_device_n takes n from the caller and stays asynchronous)
    int n0 = 0;
    {
        std::lock_guard<std::mutex> lk(E.mu);
        DEV0();
        cudaStream_t st = stream ? (cudaStream_t)stream : dv.stream;
        CK(cudaMemcpyAsync(&n0, d_fronts, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    if (n0 < 2 || n0 > CCP_MAX_N) { E.err = "device batch: unsupported n in first descriptor"; return CCP_E_ARG; }
    return ccp_frame_count_batch_device_n(d_fronts, n_fronts, d_results, stream, n0);
}

int ccp_frame_count_series(const ccp_front_desc* front, ccp_front_result* result, ccp_series_rec* series, size_t cap, size_t* len) {
    if (!E.ready) { E.err = "ccp_init not called (or no device)"; return CCP_E_NOT_INIT; }
    if (!front || !result) return CCP_E_ARG;
    std::lock_guard<std::mutex> lk(E.mu);
    DEV0();
    int rc;
    if ((rc = ensure(&dv.d_desc, &dv.desc_cap, sizeof(ccp_front_desc)))) return rc;
    if ((rc = ensure(&dv.d_res, &dv.res_cap, sizeof(ccp_front_result)))) return rc;
    size_t want = 0;
    if (front->step_log2 >= 1 && front->step_log2 <= 20 && front->L_minus > 0 && std::isfinite(front->L_minus) && front->grid >= 1 && front->grid <= CCP_MAX_GRID &&
        std::ldexp(front->L_minus, front->step_log2) <= MAX_STEPS)
        want = (size_t)ccp::step_count(front->L_minus, std::ldexp(1.0, -front->step_log2)) * (size_t)front->grid;
    if (len) *len = want;
    size_t ncap = (series && cap) ? (want < cap ? want : cap) : 0;
    if (ncap > (size_t)0x7fffffff) ncap = (size_t)0x7fffffff;
    if (ncap) { if ((rc = ensure(&dv.d_series, &dv.series_cap, ncap * sizeof(ccp_series_rec)))) return rc; }
    CK(cudaMemcpyAsync(dv.d_desc, front, sizeof(ccp_front_desc), cudaMemcpyHostToDevice, dv.stream));
    ccp_series_rec* ds = ncap ? (ccp_series_rec*)dv.d_series : nullptr;
    switch (front->n) {
        case 3: rc = launch<3>(dv, (const ccp_front_desc*)dv.d_desc, (ccp_front_result*)dv.d_res, nullptr, 1, ds, (int)ncap, dv.stream); break;
        case 4: rc = launch<4>(dv, (const ccp_front_desc*)dv.d_desc, (ccp_front_result*)dv.d_res, nullptr, 1, ds, (int)ncap, dv.stream); break;
        default: rc = launch<2>(dv, (const ccp_front_desc*)dv.d_desc, (ccp_front_result*)dv.d_res, nullptr, 1, ds, (int)ncap, dv.stream); break;
    }
    if (rc) return rc;
    CK(cudaMemcpyAsync(result, dv.d_res, sizeof(ccp_front_result), cudaMemcpyDeviceToHost, dv.stream));
    if (ncap) CK(cudaMemcpyAsync(series, dv.d_series, ncap * sizeof(ccp_series_rec), cudaMemcpyDeviceToHost, dv.stream));
    CK(cudaStreamSynchronize(dv.stream));
    return 0;
}

int ccp_frame_count_subdivided(const ccp_front_desc* fronts, size_t n_fronts, int parts, int which, ccp_subdiv_result* results) {
    if (!E.ready) { E.err = "ccp_init not called (or no device)"; return CCP_E_NOT_INIT; }
    if (n_fronts == 0) return 0;
    if (!fronts || !results || parts < 1 || parts > 64 || (which != CCP_SUBDIV_UXY && which != CCP_SUBDIV_EU_ERR)) { E.err = "ccp_frame_count_subdivided: bad argument"; return CCP_E_ARG; }
    const int n = fronts[0].n;
    if (n < 2 || n > CCP_MAX_N) { E.err = "ccp_frame_count_subdivided: unsupported n"; return CCP_E_ARG; }
    for (size_t i = 1; i < n_fronts; i++) if (fronts[i].n != n) { E.err = "ccp_frame_count_subdivided: n must be uniform over the batch"; return CCP_E_ARG; }
    long long total = 1;
    for (int q = 0; q < n; q++) total *= parts;
    if (total > (1ll << 17)) { E.err = "ccp_frame_count_subdivided: more than 2^17 sub-boxes per front"; return CCP_E_ARG; }
    std::lock_guard<std::mutex> lk(E.mu);
    DEV0();
    const size_t chunk = std::max<size_t>(1, (size_t)((1ll << 17) / total));          // fronts per pass
    int rc;
    if ((rc = ensure(&dv.d_desc, &dv.desc_cap, std::min(chunk, n_fronts) * sizeof(ccp_front_desc)))) return rc;
    if ((rc = ensure(&dv.d_sub, &dv.sub_cap, std::min(chunk, n_fronts) * (size_t)total * sizeof(ccp_front_desc)))) return rc;
    if ((rc = ensure(&dv.d_subres, &dv.subres_cap, std::min(chunk, n_fronts) * (size_t)total * sizeof(ccp_front_result)))) return rc;
    if ((rc = ensure(&dv.d_hull, &dv.hull_cap, std::min(chunk, n_fronts) * sizeof(ccp_subdiv_result)))) return rc;
    float ms_total = 0.f;
    for (size_t at = 0; at < n_fronts; at += chunk) {
        const size_t m = std::min(chunk, n_fronts - at);
        const long long nsub = (long long)m * total;
        CK(cudaMemcpyAsync(dv.d_desc, fronts + at, m * sizeof(ccp_front_desc), cudaMemcpyHostToDevice, dv.stream));
        CK(cudaEventRecord(dv.ev0, dv.stream));
        ccp::expand_kernel<<<(unsigned)((nsub + 127) / 128), 128, 0, dv.stream>>>((const ccp_front_desc*)dv.d_desc, (int)m, parts, which, (int)total, (ccp_front_desc*)dv.d_sub);
        E.launches++;
        CK(cudaGetLastError());
        if ((rc = launch_n(dv, n, (const ccp_front_desc*)dv.d_sub, (ccp_front_result*)dv.d_subres, nullptr, (int)nsub, dv.stream))) return rc;
        ccp::reduce_kernel<<<(unsigned)m, ccp::SUBDIV_NT, 0, dv.stream>>>((const ccp_front_result*)dv.d_subres, (int)total, (ccp_subdiv_result*)dv.d_hull);
        E.launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(dv.ev1, dv.stream));
        CK(cudaMemcpyAsync(results + at, dv.d_hull, m * sizeof(ccp_subdiv_result), cudaMemcpyDeviceToHost, dv.stream));
        CK(cudaStreamSynchronize(dv.stream));
        float ms = 0.f; CK(cudaEventElapsedTime(&ms, dv.ev0, dv.ev1));
        ms_total += ms;
    }
    E.last_ms = ms_total;
    return 0;
}

uint64_t ccp_kernel_launches(void) { return E.launches; }
float ccp_last_kernel_ms(void) { return E.last_ms; }

int ccp_measure_fp64_peak(int iters, double* tflops) {
    if (!E.ready) { E.err = "ccp_init not called (or no device)"; return CCP_E_NOT_INIT; }
    if (!tflops) return CCP_E_ARG;
    if (iters < 1) iters = 4096;
    std::lock_guard<std::mutex> lk(E.mu);
    DEV0();
    const int threads = 256, blocks = dv.sm_count * 8;
    double* d_out = nullptr;
    CK(cudaMalloc(&d_out, sizeof(double) * threads * blocks));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(dv.ev0, dv.stream));
        dfma_peak_kernel<<<blocks, threads, 0, dv.stream>>>(d_out, iters, 0.999999, 1e-7);
        CK(cudaEventRecord(dv.ev1, dv.stream));
        CK(cudaStreamSynchronize(dv.stream));
        float ms = 0; CK(cudaEventElapsedTime(&ms, dv.ev0, dv.ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(d_out);
    const double flop = 2.0 * 64.0 * (double)iters * threads * blocks;     // 64 DFMA per thread per iteration
    *tflops = flop / (best * 1e-3) / 1e12;
    return 0;
}

// FP64 work of one front under the implemented algorithm: DFMA = 2 flop, DADD/DMUL = 1.
// localManifold::boundDFU for a batch of calls: host buffers in, host buffers out
int ccp_bound_dfu_batch(const ccp_dfu_desc* descs, size_t n_descs, ccp_dfu_result* results) {
    std::lock_guard<std::mutex> lk(E.mu);
    if (!E.ready) { E.err = "ccp_init has not been called (no CPU fallback)"; return CCP_E_NOT_INIT; }
    if (!descs || !results) { E.err = "null argument"; return CCP_E_ARG; }
    if (n_descs == 0) return 0;
    DEV0();
    int rc;                                                                                   // the engine's staging buffers are reused
    if ((rc = ensure(&dv.d_desc, &dv.desc_cap, n_descs * sizeof(ccp_dfu_desc)))) return rc;
    if ((rc = ensure(&dv.d_res, &dv.res_cap, n_descs * sizeof(ccp_dfu_result)))) return rc;
    void* d_in = dv.d_desc; void* d_out = dv.d_res;
    CK(cudaMemcpyAsync(d_in, descs, n_descs * sizeof(ccp_dfu_desc), cudaMemcpyHostToDevice, dv.stream));
    CK(cudaMemsetAsync(d_out, 0, n_descs * sizeof(ccp_dfu_result), dv.stream));               // unused matrix entries (n < CCP_MAX_N) stay zero
    const int grid = (int)(n_descs < (size_t)(8 * dv.sm_count) ? n_descs : (size_t)(8 * dv.sm_count));
    bool has[CCP_MAX_N + 1] = {false, false, false, false, false};
    for (size_t i = 0; i < n_descs; i++) if (descs[i].n >= 2 && descs[i].n <= CCP_MAX_N) has[descs[i].n] = true;
    CK(cudaEventRecord(dv.ev0, dv.stream));
    if (has[2]) { ccp::dfu_kernel<2><<<grid, ccp::DFU_NT, sizeof(ccp::DfuSmem<2>), dv.stream>>>((const ccp_dfu_desc*)d_in, (ccp_dfu_result*)d_out, (int)n_descs); E.launches++; }
    if (has[3]) { ccp::dfu_kernel<3><<<grid, ccp::DFU_NT, sizeof(ccp::DfuSmem<3>), dv.stream>>>((const ccp_dfu_desc*)d_in, (ccp_dfu_result*)d_out, (int)n_descs); E.launches++; }
    if (has[4]) { ccp::dfu_kernel<4><<<grid, ccp::DFU_NT, sizeof(ccp::DfuSmem<4>), dv.stream>>>((const ccp_dfu_desc*)d_in, (ccp_dfu_result*)d_out, (int)n_descs); E.launches++; }
    CK(cudaGetLastError());
    CK(cudaEventRecord(dv.ev1, dv.stream));
    CK(cudaMemcpyAsync(results, d_out, n_descs * sizeof(ccp_dfu_result), cudaMemcpyDeviceToHost, dv.stream));
    CK(cudaStreamSynchronize(dv.stream));
    CK(cudaEventElapsedTime(&E.last_ms, dv.ev0, dv.ev1));
    for (size_t i = 0; i < n_descs; i++) if (descs[i].n < 2 || descs[i].n > CCP_MAX_N) { std::memset(&results[i], 0, sizeof(ccp_dfu_result)); results[i].status = CCP_ST_BAD_DESC; }
    return 0;
}

double ccp_front_flops(int n, int order, int grid, int steps) {
    // algorithmic FP64 work of one front (DFMA = 2 flop, DADD/DMUL = 1), DESIGN.md section 5
    const double p = order, g = grid, N = n, D = 2.0 * n;
    const double P2 = p * (p + 1) / 2;                                      // sum_{k<p} (k+1) Cauchy terms per product
    const double products = (5 * N - 3) + D * (3 * N - 2);                  // centre phi tables (g, q, g*w) + centre Psi tables
    const double terms = products * P2 + (p + 1) * N * N * D;               // + frame tables
    const double dfma_cauchy = 4 * terms;
    // Horner: point + range evaluation of the frame on the grid, y and Jc; derivative enclosures are formed on demand (not counted)
    const double horner_fma = 2 * ((g + 1) * N * N * p + g * N * N * p + (D + D * D) * p);
    // second variation over the hull (closed form to h^4): (mid,t) dot-product terms of the stages X1, X2a, X2b, X3, X4 + Horner in h
    const double sv_terms = (D + 1) * (3 * N * N + 10 * N * N + 12 * N * N + 6 * N * N * N) + D * D * (D + 1);
    const double sv_fma = 4 * sv_terms + (D + 1) * N * N * 48;
    // set update: fused interval*point accumulations (2 DFMA per term) and (mid,t) dot products (4 DFMA per term); frame at step start
    const double su_fma = 2 * (D * D * N * 2 * D + D * D * D + D * N * 2 * D + D * N * D) + 4 * (D * N * 3 * D + 2 * D * D + D * N * N);
    const double conv = 10 * ((4 * N - 1) + 2 * N * D + (2 * N - 1)) * p      // iv2mt / mt2iv / divisions per stored table coefficient
                      + 10 * (D + 1) * (4 * N + 12 * N * N) + 10 * 6 * D * D;  // ... and per stored second-variation / update quantity
    const double det3 = (N == 2 ? 17 : N == 3 ? 82 : 360);
    const double dets = (2 * g + 1) * det3;                                 // point + range determinants; Jacobi derivative on demand
    const double per_step = 2 * (dfma_cauchy + horner_fma + sv_fma + su_fma) + conv + dets;
    return per_step * steps;
}

} // extern "C"
