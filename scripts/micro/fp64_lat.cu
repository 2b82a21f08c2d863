// FP64 latency / throughput microbenchmark for one warp and several warps per SM (B200).
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
    double x[CH];
    for (int c = 0; c < CH; c++) x[c] = threadIdx.x * 1e-9 + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) x[c] = __fma_ru(x[c], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int c = 0; c < CH; c++) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void minmax_chain(double* out, long long* cyc, int iters, double a) {
    double x = threadIdx.x * 1e-9, y = 1.0 + x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { x = (y < x) ? y : x * a; y = (x < y) ? y : x; }   // DSETP + FSEL chains
    long long t1 = clock64();
    out[threadIdx.x] = x + y;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lds_chain(double* out, long long* cyc, int iters) {
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 17 + 5) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) p = idx[p];
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    long long h; int iters = 4096;
#define RUN(CH, BL, TH) chain<CH><<<BL, TH>>>(out, cyc, iters, 0.999999, 1e-7); cudaDeviceSynchronize(); chain<CH><<<BL, TH>>>(out, cyc, iters, 0.999999, 1e-7); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("chains=%d blocks=%d threads=%d : %.2f cycles per DFMA-per-chain-step, %.2f cycles/DFMA warp-instr (per warp)\n", CH, BL, TH, (double)h / iters, (double)h / iters / CH);
    RUN(1, 1, 32) RUN(2, 1, 32) RUN(4, 1, 32) RUN(8, 1, 32) RUN(16, 1, 32)
    RUN(1, 1, 128) RUN(4, 1, 128) RUN(8, 1, 128) RUN(4, 1, 256) RUN(4, 1, 512) RUN(8, 1, 1024)
    minmax_chain<<<1, 32>>>(out, cyc, iters, 0.999); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("minmax dependent pair: %.2f cycles/iter\n", (double)h / iters);
    lds_chain<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("LDS dependent: %.2f cycles/load\n", (double)h / iters);
    return 0;
}
