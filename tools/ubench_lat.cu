// Dependent-issue latencies of the FP64 instructions the diagonal-block kernel chains (developer tool).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(double* out, long long* cyc, double seed) {
    double a = seed + threadIdx.x, b = 1.0 + 1e-9 * threadIdx.x, c0 = 0, c1 = 0, d0 = 0, d1 = 0;
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = seed;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) dmma(c0, c1, a, b);                       // dependent DMMA chain
            if (MODE == 1) { c0 = fma(c0, b, a); }                   // dependent DFMA chain
            if (MODE == 2) { c0 = rsqrt(c0 + a); }                   // dependent rsqrt chain
            if (MODE == 3) { dmma(c0, c1, a, b); dmma(d0, d1, a, b); }  // 2 independent chains
            if (MODE == 4) { c0 = sm[((int)c0 + j) & 63] + a; }      // LDS dependent chain (+1 DADD)
            if (MODE == 5) { c0 = __shfl_sync(0xffffffffu, c0, (threadIdx.x + 1) & 31) + a; }
            if (MODE == 6) { c0 = 1.0 / (c0 + a); }
            if (MODE == 7) { c0 = sqrt(c0 + a); }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1 + d0 + d1;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1024); cudaMalloc(&cyc, 8);
    const char* names[] = {"dmma dep", "dfma dep", "rsqrt(+dadd) dep", "dmma x2 indep", "lds+dadd dep", "shfl64+dadd dep", "rcp(+dadd) dep", "sqrt(+dadd) dep"};
    for (int m = 0; m < 8; ++m) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (m) {
                case 0: k<0><<<1, 32>>>(out, cyc, 1.5); break; case 1: k<1><<<1, 32>>>(out, cyc, 1.5); break;
                case 2: k<2><<<1, 32>>>(out, cyc, 1.5); break; case 3: k<3><<<1, 32>>>(out, cyc, 1.5); break;
                case 4: k<4><<<1, 32>>>(out, cyc, 1.5); break; case 5: k<5><<<1, 32>>>(out, cyc, 1.5); break;
                case 6: k<6><<<1, 32>>>(out, cyc, 1.5); break; case 7: k<7><<<1, 32>>>(out, cyc, 1.5); break;
            }
            cudaDeviceSynchronize();
        }
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-20s %7.1f cycles per iteration\n", names[m], h / 512.0);
    }
    return 0;
}
