// TEST INFRASTRUCTURE ONLY (never linked into the product): serves the C-ABI entry points that tests/cpp/manifold_gpu.cpp and driver_gpu.cpp need from
// the CPU oracle twins of the kernels (oracle/dfu.hpp, oracle/c2.hpp via liboracle.so), so that the host mirrors around boundDFU can be
// exercised on a machine without a GPU.  The CUDA kernel is bit-identical to this oracle (tests/test_bound_dfu.py, -m gpu).
#include "../../include/ccp.h"
#include <cstddef>
extern "C" int orc_bound_dfu(const ccp_dfu_desc* d, size_t n, ccp_dfu_result* res);
extern "C" int orc_run_batch(const ccp_front_desc* d, size_t n_fronts, int mode, ccp_front_result* res, int threads, double* seconds);
extern "C" {
int ccp_device_count(void) { return 1; }
int ccp_init(int) { return 0; }
const char* ccp_last_error(void) { return "oracle backend"; }
int ccp_bound_dfu_batch(const ccp_dfu_desc* d, size_t n, ccp_dfu_result* r) { return orc_bound_dfu(d, n, r); }
int ccp_frame_count_batch(const ccp_front_desc* d, size_t n, ccp_front_result* r) { double s; return orc_run_batch(d, n, /*mode c2 = kernel twin*/ 2, r, 0, &s); }
int ccp_frame_count_batch_multi(const ccp_front_desc* d, size_t n, ccp_front_result* r) { return ccp_frame_count_batch(d, n, r); }
int ccp_frame_count_series(const ccp_front_desc*, ccp_front_result*, ccp_series_rec*, size_t, size_t*) { return CCP_E_NO_DEVICE; }
}
