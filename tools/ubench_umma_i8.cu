// Microbenchmark + correctness probe for the integer tensor-core path of DESIGN.md section 10 (item 1):
// tcgen05.mma kind::i8 (int8 x int8 -> int32 in TMEM) on sm_100a, operands in shared memory in the canonical
// K-major no-swizzle ("interleave") layout, one CTA per SM.  Hand-written PTX, no CUTLASS.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_umma_i8 tools/ubench_umma_i8.cu
//   tools/ubench_umma_i8 [reps=2000]
//
// Prints (1) whether D = A B^T matches the CPU for a random 128 x N x K problem and (2) the sustained rate of
// back-to-back MMAs re-using the same operands (issue-rate peak of the int8 pipe), per SM and for the chip.
// Layout (units of 16 bytes): element (row, kbyte) of an operand with R rows lives at
//   slab = kbyte / 16;  offset = slab * (R * 16) + row * 16 + kbyte % 16
// i.e. 8-row x 16-byte core matrices, consecutive along the rows (SBO = 128 B), slabs along K (LBO = R * 16 B).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor), SWIZZLE_NONE, K-major
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);             // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;   // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;   // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                             // version = 1 (Blackwell), bits [46,48)
    return d;                                           // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// instruction descriptor (InstrDescriptor): S32 accumulate, signed int8 A and B, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int N>
__global__ void __launch_bounds__(128, 1)
k_umma_i8(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int K, int32_t* __restrict__ D, int reps,
          long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nslab = K / 16;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)nslab * 128 * 16;
    for (int idx = tid; idx < 128 * nslab; idx += 128) {
        const int row = idx % 128, slab = idx / 128;
        *reinterpret_cast<int4*>(sA + (size_t)slab * 2048 + row * 16) =
            *reinterpret_cast<const int4*>(A + (size_t)row * K + slab * 16);
    }
    for (int idx = tid; idx < N * nslab; idx += 128) {
        const int row = idx % N, slab = idx / N;
        *reinterpret_cast<int4*>(sB + (size_t)slab * (N * 16) + row * 16) =
            *reinterpret_cast<const int4*>(B + (size_t)row * K + slab * 16);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy stores -> visible to the tensor core (async proxy)
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)(N < 32 ? 32 : N)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    long long t0 = 0, t1 = 0;
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = umma_idesc_i8(128, N);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        t0 = clock64();
        for (int rep = 0; rep < reps; ++rep) {
            for (int k = 0; k < K / 32; ++k) {
                const uint64_t ad = umma_desc(a0 + k * 2 * 2048, 2048, 128);
                const uint64_t bd = umma_desc(b0 + k * 2 * (N * 16), N * 16, 128);
                const uint32_t accumulate = (rep | k) != 0;
                asm volatile(
                    "{\n\t"
                    ".reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
                    "}\n" ::"r"(tmem),
                    "l"(ad), "l"(bd), "r"(idesc), "r"(accumulate), "r"(0u));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                     : "memory");
    }
    {   // everybody waits for the MMAs (phase 0 of the barrier)
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t"
                "}\n"
                : "=r"(done)
                : "r"(smem_u32(&bar)), "r"(0u)
                : "memory");
        }
    }
    if (warp == 1 && lane == 0) {
        t1 = clock64();
        if (cycles) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // epilogue: warp w owns TMEM lanes 32 w .. 32 w + 31 (= rows of D), 8 columns per load
    if (blockIdx.x == 0 && D) {
        const int row = warp * 32 + lane;
        for (int c0 = 0; c0 < N; c0 += 8) {
            uint32_t r[8];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int u = 0; u < 8; ++u) D[(size_t)row * N + c0 + u] = (int32_t)r[u];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)(N < 32 ? 32 : N)));
}

template <int N>
int run(int K, int reps) {
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    std::vector<int8_t> hA((size_t)128 * K), hB((size_t)N * K);
    srand(1234 + N);
    for (auto& v : hA) v = (int8_t)(rand() % 255 - 127);
    for (auto& v : hB) v = (int8_t)(rand() % 255 - 127);
    int8_t *dA, *dB;
    int32_t* dD;
    long long* dC;
    CK(cudaMalloc(&dA, hA.size()));
    CK(cudaMalloc(&dB, hB.size()));
    CK(cudaMalloc(&dD, sizeof(int32_t) * 128 * N));
    CK(cudaMalloc(&dC, sizeof(long long) * sms));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(K / 16) * (2048 + N * 16);
    CK(cudaFuncSetAttribute(k_umma_i8<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // (1) correctness: one pass
    k_umma_i8<N><<<1, 128, smem>>>(dA, dB, K, dD, 1, nullptr);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> hD((size_t)128 * N);
    CK(cudaMemcpy(hD.data(), dD, sizeof(int32_t) * hD.size(), cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < N; ++j) {
            int32_t s = 0;
            for (int k = 0; k < K; ++k) s += (int32_t)hA[(size_t)i * K + k] * (int32_t)hB[(size_t)j * K + k];
            bad += s != hD[(size_t)i * N + j];
        }
    // (2) rate: every SM, reps passes over the same operands
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k_umma_i8<N><<<sms, 128, smem>>>(dA, dB, K, nullptr, 10, dC);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_umma_i8<N><<<sms, 128, smem>>>(dA, dB, K, nullptr, reps, dC);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> hC(sms);
    CK(cudaMemcpy(hC.data(), dC, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double cyc = 0;
    for (auto c : hC) cyc += (double)c;
    cyc /= sms;
    const double mmas = (double)reps * (K / 32);
    const double ops_per_mma = 2.0 * 128 * N * 32;
    printf("{\"N\": %d, \"K\": %d, \"mismatches\": %lld, \"cycles_per_mma\": %.1f, \"macs_per_clk_per_sm\": %.0f, "
           "\"tops_chip_by_event\": %.1f, \"sms\": %d}\n",
           N, K, bad, cyc / mmas, 128.0 * N * 32 / (cyc / mmas), mmas * ops_per_mma * sms / (ms * 1e-3) / 1e12, sms);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
    return bad != 0;
}

int main(int argc, char** argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 2000;
    int rc = 0;
    rc |= run<64>(128, reps);
    rc |= run<128>(256, reps);
    rc |= run<256>(256, reps);
    return rc;
}
