// oracle/ival.hpp -- TEST INFRASTRUCTURE (CPU oracle).  Not part of the product path.
//
// Host interval arithmetic with outward rounding, restating what the reference gets from
// CAPD's `interval` (capd::intervals::Interval<double>; call sites: every arithmetic
// expression in propagateManifold.cpp / utils.cpp / topFrame.cpp).  [CAPD-recollection]:
// closed [lo,hi] doubles, every operation outward rounded, comparisons have "certainly"
// semantics (a>b <=> a.lo > b.hi).
//
// Convention: the FPU is kept in FE_UPWARD for the whole oracle run (RoundUpGuard); lower
// bounds are obtained with the negation identity rd(x op y) = -ru((-x) op' y).  Compile with
// -frounding-math -ffp-contract=off (see oracle/Makefile); std::fma gives the correctly
// rounded fused operation, so every primitive is bit-identical to the CUDA intrinsic of the
// same name (__dadd_ru, __dmul_rd, __fma_ru, __ddiv_rd, ...).
#pragma once
#include <cfenv>
#include <cmath>
#include <algorithm>
#include <limits>

namespace orc {

struct RoundUpGuard {
    int old;
    RoundUpGuard() : old(fegetround()) { fesetround(FE_UPWARD); }
    ~RoundUpGuard() { fesetround(old); }
};

struct RoundNearestGuard {          // for the few round-to-nearest operations of the (mid,t) scalings
    int old;
    RoundNearestGuard() : old(fegetround()) { fesetround(FE_TONEAREST); }
    ~RoundNearestGuard() { fesetround(old); }
};

// ---- directed primitives (FE_UPWARD assumed) ----
// `volatile`-free: -frounding-math forbids the re-association that would break these.
static inline double add_ru(double a, double b) { return a + b; }
static inline double add_rd(double a, double b) { return -((-a) - b); }
static inline double sub_ru(double a, double b) { return a - b; }
static inline double sub_rd(double a, double b) { return -(b - a); }
static inline double mul_ru(double a, double b) { return a * b; }
static inline double mul_rd(double a, double b) { return -((-a) * b); }
static inline double div_ru(double a, double b) { return a / b; }
static inline double div_rd(double a, double b) { return -((-a) / b); }
static inline double fma_ru(double a, double b, double c) { return std::fma(a, b, c); }
// The empty asm is an optimisation barrier: without it GCC folds -fma(-a,b,-c) into fma(a,b,c) (one vfmadd)
// even under -frounding-math, which silently rounds the LOWER bound UP.  Caught by the bit-for-bit
// comparison with __fma_rd on the GPU (tests/test_gpu_primitives.py).
static inline double fma_rd(double a, double b, double c) {
    double t = std::fma(-a, b, -c);
    __asm__ volatile("" : "+x"(t));
    return -t;
}
static inline double sqrt_ru(double a) { return std::sqrt(a); }

struct ival {
    double lo, hi;
    ival() : lo(0.0), hi(0.0) {}
    ival(double x) : lo(x), hi(x) {}
    ival(double l, double h) : lo(l), hi(h) {}
};

static inline ival operator+(const ival& a, const ival& b) { return ival(add_rd(a.lo, b.lo), add_ru(a.hi, b.hi)); }
static inline ival operator-(const ival& a, const ival& b) { return ival(sub_rd(a.lo, b.hi), sub_ru(a.hi, b.lo)); }
static inline ival operator-(const ival& a) { return ival(-a.hi, -a.lo); }

// general-sign product: 4 candidates per bound (CAPD does sign-case analysis; the result is
// the same set of end points).
static inline ival operator*(const ival& a, const ival& b) {
    double l = std::min(std::min(mul_rd(a.lo, b.lo), mul_rd(a.lo, b.hi)),
                        std::min(mul_rd(a.hi, b.lo), mul_rd(a.hi, b.hi)));
    double h = std::max(std::max(mul_ru(a.lo, b.lo), mul_ru(a.lo, b.hi)),
                        std::max(mul_ru(a.hi, b.lo), mul_ru(a.hi, b.hi)));
    return ival(l, h);
}
static inline ival sqr(const ival& a) {
    if (a.lo >= 0) return ival(mul_rd(a.lo, a.lo), mul_ru(a.hi, a.hi));
    if (a.hi <= 0) return ival(mul_rd(a.hi, a.hi), mul_ru(a.lo, a.lo));
    return ival(0.0, std::max(mul_ru(a.lo, a.lo), mul_ru(a.hi, a.hi)));
}
// division by an interval not containing zero
static inline ival operator/(const ival& a, const ival& b) {
    double l = std::min(std::min(div_rd(a.lo, b.lo), div_rd(a.lo, b.hi)),
                        std::min(div_rd(a.hi, b.lo), div_rd(a.hi, b.hi)));
    double h = std::max(std::max(div_ru(a.lo, b.lo), div_ru(a.lo, b.hi)),
                        std::max(div_ru(a.hi, b.lo), div_ru(a.hi, b.hi)));
    return ival(l, h);
}
static inline ival& operator+=(ival& a, const ival& b) { a = a + b; return a; }
static inline ival& operator-=(ival& a, const ival& b) { a = a - b; return a; }
static inline ival& operator*=(ival& a, const ival& b) { a = a * b; return a; }

static inline ival hull(const ival& a, const ival& b) { return ival(std::min(a.lo, b.lo), std::max(a.hi, b.hi)); }
static inline bool intersect(const ival& a, const ival& b, ival& out) {
    double l = std::max(a.lo, b.lo), h = std::min(a.hi, b.hi);
    if (l > h) return false;
    out = ival(l, h);
    return true;
}
static inline double mid(const ival& a) {             // any point of a is fine; keep it inside
    double m = 0.5 * a.lo + 0.5 * a.hi;
    if (m < a.lo) m = a.lo;
    if (m > a.hi) m = a.hi;
    return m;
}
static inline double mag(const ival& a) { return std::max(std::fabs(a.lo), std::fabs(a.hi)); }
static inline double width(const ival& a) { return sub_ru(a.hi, a.lo); }
static inline bool certainly_pos(const ival& a) { return a.lo > 0; }
static inline bool certainly_neg(const ival& a) { return a.hi < 0; }
static inline bool definite(const ival& a) { return a.lo > 0 || a.hi < 0; }
static inline bool subset(const ival& a, const ival& b) { return b.lo <= a.lo && a.hi <= b.hi; }
static inline bool subset_interior(const ival& a, const ival& b) { return b.lo < a.lo && a.hi < b.hi; }
static inline bool finite(const ival& a) { return std::isfinite(a.lo) && std::isfinite(a.hi); }
static inline ival sym(double r) { return ival(-r, r); }

} // namespace orc
