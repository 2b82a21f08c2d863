// Issue cost of FP64 compare/select vs integer sign test, one warp and 4 warps per SMSP (B200).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, long long* cyc, int iters, double tl, double th) {
    double x[8];
    for (int c = 0; c < 8; c++) x[c] = (threadIdx.x & 1 ? -1.0 : 1.0) * (1e-3 + c);
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            if (MODE == 0) x[c] = __fma_rd(x[c], (x[c] >= 0) ? tl : th, 1e-9);                       // DSETP + FSEL*2 + DFMA
            if (MODE == 1) x[c] = __fma_rd(x[c], (__double2hiint(x[c]) >= 0) ? tl : th, 1e-9);       // ISETP + SEL*2 + DFMA
            if (MODE == 2) x[c] = __fma_rd(x[c], tl, 1e-9);                                          // DFMA only
            if (MODE == 3) x[c] = (x[c] < th) ? th * x[c] : x[c];                                    // dmax-like: DSETP + DMUL + FSEL
            if (MODE == 4) x[c] = __dadd_rd(x[c], tl);
            if (MODE == 5) x[c] = __dmul_rd(x[c], tl);
        }
    }
    long long t1 = clock64();
    double s = 0; for (int c = 0; c < 8; c++) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    long long h; int iters = 2048;
    const char* nm[] = {"DSETP+FSEL+DFMA", "ISETP+SEL+DFMA", "DFMA", "DSETP+DMUL+FSEL", "DADD", "DMUL"};
#define RUN(M, TH) k<M><<<1, TH>>>(out, cyc, iters, 0.999, 0.998); cudaDeviceSynchronize(); k<M><<<1, TH>>>(out, cyc, iters, 0.999, 0.998); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-18s threads=%4d : %.2f cycles per 8-chain step (%.2f per op group)\n", nm[M], TH, (double)h / iters, (double)h / iters / 8);
    RUN(0, 32) RUN(1, 32) RUN(2, 32) RUN(3, 32) RUN(4, 32) RUN(5, 32)
    RUN(0, 128) RUN(1, 128) RUN(2, 128) RUN(3, 128) RUN(4, 128) RUN(5, 128)
    RUN(0, 512) RUN(1, 512) RUN(2, 512) RUN(3, 512) RUN(4, 512) RUN(5, 512)
    return 0;
}
