// Variants of the point-Horner loop: what bounds a step (2 LDS.128 + 4 DFMA)?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../computing-conjugate-points_b200/csrc/ccp_interval.cuh"
using namespace ccp;
// MODE 0: as in the kernel (index k*NN+c).  1: pointer walking.  2: every lane its own column (no broadcast; table 21x64).
// 3: pointer walking + explicit one-step-ahead prefetch.  4: one entry per lane only.  5: four entries per lane, pointer walking
template <int MODE>
__global__ void k(double* out, long long* cyc, int p, double tau) {
    constexpr int NN = 9;
    __shared__ ival Fk[21 * 64];
    for (int i = threadIdx.x; i < 21 * 64; i += blockDim.x) Fk[i] = mk(1.0 / (i + 1) - 1e-12, 1.0 / (i + 1) + 1e-12);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int c0 = lane % NN, c1 = (lane + 32) % NN;
    ival a0, a1, a2, a3;
    long long t0 = clock64();
    if (MODE == 0) {
        a0 = Fk[p * NN + c0]; a1 = Fk[p * NN + c1];
#pragma unroll 4
        for (int k = p - 1; k >= 0; k--) { a0 = horner_pt(a0, tau, Fk[k * NN + c0]); a1 = horner_pt(a1, tau, Fk[k * NN + c1]); }
    } else if (MODE == 1) {
        const ival* q0 = Fk + p * NN + c0; const ival* q1 = Fk + p * NN + c1;
        a0 = *q0; a1 = *q1;
#pragma unroll 4
        for (int k = p - 1; k >= 0; k--) { q0 -= NN; q1 -= NN; a0 = horner_pt(a0, tau, *q0); a1 = horner_pt(a1, tau, *q1); }
    } else if (MODE == 2) {
        const ival* q0 = Fk + p * 64 + lane; const ival* q1 = Fk + p * 64 + 32 + lane;
        a0 = *q0; a1 = *q1;
#pragma unroll 4
        for (int k = p - 1; k >= 0; k--) { q0 -= 64; q1 -= 64; a0 = horner_pt(a0, tau, *q0); a1 = horner_pt(a1, tau, *q1); }
    } else if (MODE == 3) {
        const ival* q0 = Fk + p * NN + c0; const ival* q1 = Fk + p * NN + c1;
        a0 = *q0; a1 = *q1; q0 -= NN; q1 -= NN;
        ival n0 = *q0, n1 = *q1;
#pragma unroll 4
        for (int k = p - 1; k >= 0; k--) { const ival d0 = n0, d1 = n1; if (k > 0) { q0 -= NN; q1 -= NN; n0 = *q0; n1 = *q1; } a0 = horner_pt(a0, tau, d0); a1 = horner_pt(a1, tau, d1); }
    } else if (MODE == 4) {
        const ival* q0 = Fk + p * NN + c0;
        a0 = *q0; a1 = a0;
#pragma unroll 4
        for (int k = p - 1; k >= 0; k--) { q0 -= NN; a0 = horner_pt(a0, tau, *q0); }
    } else {
        const int c2 = (lane + 64) % NN, c3 = (lane + 96) % NN;
        const ival* q0 = Fk + p * NN + c0; const ival* q1 = Fk + p * NN + c1; const ival* q2 = Fk + p * NN + c2; const ival* q3 = Fk + p * NN + c3;
        a0 = *q0; a1 = *q1; a2 = *q2; a3 = *q3;
#pragma unroll 4
        for (int k = p - 1; k >= 0; k--) { q0 -= NN; q1 -= NN; q2 -= NN; q3 -= NN; a0 = horner_pt(a0, tau, *q0); a1 = horner_pt(a1, tau, *q1); a2 = horner_pt(a2, tau, *q2); a3 = horner_pt(a3, tau, *q3); }
        a0 = iadd(a0, a2); a1 = iadd(a1, a3);
    }
    long long t1 = clock64();
    out[threadIdx.x] = a0.lo + a1.hi;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
    long long h;
    const char* nm[] = {"index k*NN+c", "pointer walk", "no broadcast", "explicit prefetch", "one entry", "four entries"};
#define RUN(M) k<M><<<1, 64>>>(out, cyc, 20, 0.03); cudaDeviceSynchronize(); k<M><<<1, 64>>>(out, cyc, 20, 0.03); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-18s: %lld cycles for 20 steps (%.1f per step)\n", nm[M], h, (double)h / 20);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    return 0;
}
