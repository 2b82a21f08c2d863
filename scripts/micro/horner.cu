// Isolated copy of the grid-evaluation loops of the front kernel (point / range Horner over Fk in shared memory).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../computing-conjugate-points_b200/csrc/ccp_interval.cuh"
using namespace ccp;
template <int MODE, int UNR>
__global__ void k(double* out, long long* cyc, int p, int g, double h) {
    constexpr int NN = 9;
    __shared__ ival Fk[21 * NN];
    __shared__ double gp[16];
    __shared__ ival EV[(2 * 15 + 1) * NN];
    for (int i = threadIdx.x; i < 21 * NN; i += blockDim.x) Fk[i] = mk(1.0 / (i + 1) - 1e-12, 1.0 / (i + 1) + 1e-12);
    if (threadIdx.x <= g) gp[threadIdx.x] = h * threadIdx.x / g;
    __syncthreads();
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int E = g * NN;
    const ival* Fk0 = Fk;
    long long t0 = clock64();
    for (int e0 = lane; e0 < E; e0 += 64) {
        const int e1 = e0 + 32; const bool v1 = e1 < E; const int ee1 = v1 ? e1 : e0;
        const int c0 = e0 % NN, c1 = ee1 % NN, g0 = e0 / NN, g1 = ee1 / NN;
        ival a0 = Fk0[p * NN + c0], a1 = Fk0[p * NN + c1];
        if (MODE == 0) {
            const double t0 = gp[g0 + 1], t1 = gp[g1 + 1];
#pragma unroll UNR
            for (int k = p - 1; k >= 0; k--) { a0 = horner_pt(a0, t0, Fk0[k * NN + c0]); a1 = horner_pt(a1, t1, Fk0[k * NN + c1]); }
            EV[NN + e0] = a0; if (v1) EV[NN + e1] = a1;
        } else {
            const double l0 = gp[g0], h0 = gp[g0 + 1], l1 = gp[g1], h1 = gp[g1 + 1];
#pragma unroll UNR
            for (int k = p - 1; k >= 0; k--) { a0 = horner_rg(a0, l0, h0, Fk0[k * NN + c0]); a1 = horner_rg(a1, l1, h1, Fk0[k * NN + c1]); }
            EV[(g + 1) * NN + e0] = a0; if (v1) EV[(g + 1) * NN + e1] = a1;
        }
    }
    long long t1 = clock64();
    __syncthreads();
    out[threadIdx.x] = EV[NN + threadIdx.x].lo + EV[(g + 1) * NN + threadIdx.x].hi;
    if (lane == 0) cyc[wq] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
    long long h[2];
#define RUN(M, U) k<M, U><<<1, 64>>>(out, cyc, 20, 14, 0.03125); cudaDeviceSynchronize(); k<M, U><<<1, 64>>>(out, cyc, 20, 14, 0.03125); cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost); \
    printf("mode %d unroll %d: warp0 %lld warp1 %lld cycles for 2 rounds x 20 steps (%.1f per step)\n", M, U, h[0], h[1], (double)h[0] / 40);
    RUN(0, 1) RUN(0, 4) RUN(1, 1) RUN(1, 4)
    return 0;
}
