// Micro-benchmarks that size the FP64 roofline on this B200: DMMA shapes vs DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64 ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int NACC>
__global__ void k_dmma884(double* out, int iters) {
    double c[NACC][2];
    #pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
    #pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma16816(double* out, int iters) {
    double c[NACC][4];
    #pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0; }
    double a[8], b[4];
    #pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    #pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 2e-3 + i;
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0;
    #pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma1688(double* out, int iters) {
    double c[NACC][4];
    #pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0; }
    double a[4], b[2];
    #pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = threadIdx.x * 1e-3 + i;
    #pragma unroll
    for (int i = 0; i < 2; ++i) b[i] = threadIdx.x * 2e-3 + i;
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
        }
    }
    double s = 0;
    #pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters) {
    double c[NACC];
    #pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = threadIdx.x * 1e-3, b = 1.0000001;
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], b, a);
    }
    double s = 0;
    #pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        for (int bpsm : {1, 2}) {
            int grid = sms * bpsm, block = warps * 32;
            if (warps * bpsm > 64) continue;
            double nw = (double)grid * warps;
            float ms;
            ms = timeit([&] { k_dmma884<8><<<grid, block>>>(out, iters); });
            printf("dmma884   acc8  warps/blk %2d blk/sm %d: %8.2f TFLOP/s\n", warps, bpsm, nw * iters * 8 * 512.0 / ms / 1e9);
            ms = timeit([&] { k_dmma884<16><<<grid, block>>>(out, iters); });
            printf("dmma884   acc16 warps/blk %2d blk/sm %d: %8.2f TFLOP/s\n", warps, bpsm, nw * iters * 16 * 512.0 / ms / 1e9);
            ms = timeit([&] { k_dmma1688<8><<<grid, block>>>(out, iters); });
            printf("dmma1688  acc8  warps/blk %2d blk/sm %d: %8.2f TFLOP/s\n", warps, bpsm, nw * iters * 8 * 2048.0 / ms / 1e9);
            ms = timeit([&] { k_dmma16816<8><<<grid, block>>>(out, iters); });
            printf("dmma16816 acc8  warps/blk %2d blk/sm %d: %8.2f TFLOP/s\n", warps, bpsm, nw * iters * 8 * 4096.0 / ms / 1e9);
            ms = timeit([&] { k_dfma<16><<<grid, block>>>(out, iters); });
            printf("dfma      acc16 warps/blk %2d blk/sm %d: %8.2f TFLOP/s\n", warps, bpsm, nw * iters * 16 * 64.0 / ms / 1e9);
        }
    }
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    return 0;
}
