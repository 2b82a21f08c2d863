// oracle/subdiv.hpp -- TEST INFRASTRUCTURE (CPU oracle).  Not part of the product path.
//
// Serial twin of csrc/ccp_subdiv.cuh (BASELINE config 4): the box of a front cut into parts^n sub-boxes with the reference's rule
// part() (utils.cpp:3-10; index rule of getSubdivision, utils.cpp:309-321), each sub-box run through an oracle mode, counts tallied
// and every enclosure of the result record hulled.
#pragma once
#include "common.hpp"
#include <cstddef>

namespace orc {
namespace subdiv {

static inline ival mulk(const ival& a, int k) { return ival(mul_rd((double)k, a.lo), mul_ru((double)k, a.hi)); }
static inline ival divn(const ival& a, int n) { return ival(div_rd(a.lo, (double)n), div_ru(a.hi, (double)n)); }
static inline ival part(const ival& x, int k, int N) {
    const ival xl(x.lo), xr(x.hi);
    return (xl + divn(mulk(xr - xl, k), N)) + divn(x - xl, N);
}

static inline ccp_front_desc sub_desc(const ccp_front_desc& base, int j, int parts, int which) {
    ccp_front_desc d = base;
    const int n = d.n;
    int rest = j;
    for (int q = 0; q < n && q < CCP_MAX_N; q++) {
        const int k = rest % parts; rest /= parts;
        if (which == CCP_SUBDIV_UXY) {
            const ival p = part(ival(d.Uxy[q].lo, d.Uxy[q].hi), k, parts);
            d.Uxy[q].lo = p.lo; d.Uxy[q].hi = p.hi;
        } else {
            for (int c = 0; c < n; c++) {
                const ccp_interval e = d.Eu_err[q * CCP_MAX_N + c];
                const ival p = part(ival(e.lo, e.hi), k, parts);
                d.Eu_err[q * CCP_MAX_N + c].lo = p.lo; d.Eu_err[q * CCP_MAX_N + c].hi = p.hi;
            }
        }
    }
    return d;
}

static inline void reduce(const ccp_front_result* R, int total, ccp_subdiv_result& o) {
    std::memset(&o, 0, sizeof(o));
    int cert = 0, fail = 0, zmin = 0x7fffffff, zmax = -1, st = CCP_OK;
    for (int s = 0; s < total; s++) {
        const bool ok = R[s].status == CCP_OK && R[s].failure_count == 0;
        cert += ok; fail += R[s].failure_count; st = std::max(st, (int)R[s].status);
        if (ok) { zmin = std::min(zmin, (int)R[s].zero_count); zmax = std::max(zmax, (int)R[s].zero_count); }
    }
    o.status = st; o.sub_boxes = total; o.certified = cert; o.failures = fail;
    o.zero_min = cert ? zmin : -1; o.zero_max = cert ? zmax : -1;
    o.zero_count = (cert == total && zmin == zmax) ? zmin : -1;
    o.steps = R[0].steps;
    const int niv = (int)((sizeof(ccp_front_result) - offsetof(ccp_front_result, last_frame)) / sizeof(ccp_interval));
    ccp_interval* out = reinterpret_cast<ccp_interval*>(reinterpret_cast<char*>(&o) + offsetof(ccp_subdiv_result, last_frame));
    for (int i = 0; i < niv; i++) {
        double lo = INFINITY, hi = -INFINITY;
        for (int s = 0; s < total; s++) {
            const ccp_interval v = reinterpret_cast<const ccp_interval*>(reinterpret_cast<const char*>(R + s) + offsetof(ccp_front_result, last_frame))[i];
            lo = (v.lo < lo) ? v.lo : lo; hi = (hi < v.hi) ? v.hi : hi;
        }
        out[i].lo = lo; out[i].hi = hi;
    }
}

} // namespace subdiv
} // namespace orc
