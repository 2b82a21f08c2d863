// ccp_subdiv.cuh -- BASELINE config 4: the initial box of a front subdivided into parts^n sub-boxes, each propagated and counted
// independently by ccp::front_kernel, then hulled / tallied per parameter ON THE DEVICE (include/ccp.h: ccp_frame_count_subdivided).
//
//   expand_kernel : one thread per sub-box writes its descriptor = the parameter's descriptor with the subdivided coordinates replaced
//                   by their parts, the reference's rule  part(x,k,N) = x.left() + k (x.right() - x.left())/N + (x - x.left())/N
//                   (utils.cpp:3-10; index rule of getSubdivision, utils.cpp:309-321: part_list[q] = (j / N^q) % N)
//   reduce_kernel : one CTA per parameter.  Tally of the counts with warp ballots / shuffles (all sub-boxes must agree on the count and
//                   none may fail for the parameter to be certified), interval hull of every enclosure of the result records
//                   (last frame, first frame, determinants, flow map): thread i owns interval i of the record and runs over the
//                   sub-boxes, so every load of a warp is one contiguous 512-byte piece of a record.
// Serial twin: oracle/subdiv.hpp.
#pragma once
#include "ccp_interval.cuh"
#include "../../include/ccp.h"

namespace ccp {

__device__ __forceinline__ ival sub_divn(const ival& a, int n) { return mk(__ddiv_rd(a.lo, (double)n), __ddiv_ru(a.hi, (double)n)); }
__device__ __forceinline__ ival sub_part(const ival& x, int k, int N) {
    const ival xl = pt(x.lo), xr = pt(x.hi);
    return iadd(iadd(xl, sub_divn(imulk(isub(xr, xl), k), N)), sub_divn(isub(x, xl), N));
}

// sub-box j of parameter f -> d_sub[f * total + j]
__global__ void expand_kernel(const ccp_front_desc* __restrict__ base, int n_fronts, int parts, int which, int total, ccp_front_desc* __restrict__ sub) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_fronts * total) return;
    const int f = (int)(t / total), j = (int)(t % total);
    ccp_front_desc d = base[f];
    const int n = d.n;
    int rest = j;
    for (int q = 0; q < n && q < CCP_MAX_N; q++) {
        const int k = rest % parts; rest /= parts;
        if (which == CCP_SUBDIV_UXY) {                    // the x-part of getPointNbd's box (local unstable coordinates)
            const ival p = sub_part(mk(d.Uxy[q].lo, d.Uxy[q].hi), k, parts);
            d.Uxy[q].lo = p.lo; d.Uxy[q].hi = p.hi;
        } else {                                          // row q of Eu_m_Error_Final, every column the same way
            for (int c = 0; c < n; c++) {
                const ccp_interval e = d.Eu_err[q * CCP_MAX_N + c];
                const ival p = sub_part(mk(e.lo, e.hi), k, parts);
                d.Eu_err[q * CCP_MAX_N + c].lo = p.lo; d.Eu_err[q * CCP_MAX_N + c].hi = p.hi;
            }
        }
    }
    sub[t] = d;
}

constexpr int SUBDIV_NT = 256;
constexpr int SUBDIV_NIV = (int)((sizeof(ccp_front_result) - offsetof(ccp_front_result, last_frame)) / sizeof(ccp_interval));   // intervals of a record
static_assert(offsetof(ccp_front_result, last_frame) % 16 == 0 && SUBDIV_NIV <= SUBDIV_NT, "one thread per interval of the record");
static_assert(offsetof(ccp_subdiv_result, last_frame) % 16 == 0 &&
              sizeof(ccp_subdiv_result) - offsetof(ccp_subdiv_result, last_frame) == sizeof(ccp_front_result) - offsetof(ccp_front_result, last_frame),
              "the hull record carries the same enclosures as a front record");

__global__ void __launch_bounds__(SUBDIV_NT) reduce_kernel(const ccp_front_result* __restrict__ rec, int total, ccp_subdiv_result* __restrict__ out) {
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
    const ccp_front_result* R = rec + (size_t)f * total;
    __shared__ int s_cert[SUBDIV_NT / 32], s_fail[SUBDIV_NT / 32], s_zmin[SUBDIV_NT / 32], s_zmax[SUBDIV_NT / 32], s_st[SUBDIV_NT / 32];
    // ---- tally: ballots count the certified sub-boxes, shuffles reduce min / max count, worst status, failures
    int cert = 0, fail = 0, zmin = 0x7fffffff, zmax = -1, st = CCP_OK;
    for (int s0 = wq * 32; s0 < total; s0 += SUBDIV_NT) {
        const int s = s0 + lane;
        int status = CCP_OK, zc = 0, fc = 0; bool have = s < total;
        if (have) { status = R[s].status; zc = R[s].zero_count; fc = R[s].failure_count; }
        const bool ok = have && status == CCP_OK && fc == 0;
        cert += __popc(__ballot_sync(0xffffffffu, ok));
        if (have) { fail += fc; if (status > st) st = status; }
        if (ok) { zmin = min(zmin, zc); zmax = max(zmax, zc); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        fail += __shfl_xor_sync(0xffffffffu, fail, o);
        zmin = min(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
        zmax = max(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
        st = max(st, __shfl_xor_sync(0xffffffffu, st, o));
    }
    if (lane == 0) { s_cert[wq] = cert; s_fail[wq] = fail; s_zmin[wq] = zmin; s_zmax[wq] = zmax; s_st[wq] = st; }
    // ---- hull of every enclosure: thread i owns interval i of the record
    if (tid < SUBDIV_NIV) {
        double lo = INFINITY, hi = -INFINITY;
        for (int s = 0; s < total; s++) {
            const ccp_interval v = reinterpret_cast<const ccp_interval*>(reinterpret_cast<const char*>(R + s) + offsetof(ccp_front_result, last_frame))[tid];
            lo = dmin(lo, v.lo); hi = dmax(hi, v.hi);
        }
        ccp_interval* o = reinterpret_cast<ccp_interval*>(reinterpret_cast<char*>(out + f) + offsetof(ccp_subdiv_result, last_frame));
        o[tid].lo = lo; o[tid].hi = hi;
    }
    __syncthreads();
    if (tid == 0) {
        int c = 0, fl = 0, mn = 0x7fffffff, mx = -1, stw = CCP_OK;
        for (int w = 0; w < SUBDIV_NT / 32; w++) { c += s_cert[w]; fl += s_fail[w]; mn = min(mn, s_zmin[w]); mx = max(mx, s_zmax[w]); stw = max(stw, s_st[w]); }
        ccp_subdiv_result& o = out[f];
        o.status = stw; o.sub_boxes = total; o.certified = c; o.failures = fl;
        o.zero_min = c ? mn : -1; o.zero_max = c ? mx : -1;
        o.zero_count = (c == total && mn == mx) ? mn : -1;      // the parameter is certified iff EVERY sub-box is and all agree
        o.steps = R[0].steps;
    }
}

} // namespace ccp
