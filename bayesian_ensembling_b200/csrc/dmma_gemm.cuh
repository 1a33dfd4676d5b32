// FP64 tensor-core (DMMA m8n8k4) tile engine for sm_100a.
//
// One CTA (128 threads = 4 warps as 2 x 2, warp tile 64 x 32) computes a 128 x 64 tile of
//     acc[i, j] = sum_k A[i, k] * B[j, k]            ("NT": both operands K-contiguous)
// TWO such CTAs are resident per SM (248 registers x 128 threads and 96 KB of shared memory
// each): while one sits in its prologue, epilogue or a stage barrier the other keeps the FP64
// tensor pipe busy -- with one 256-thread CTA per SM the pipe measured 82% busy (ncu,
// profiles/r01c) because every bubble of the only resident CTA was a bubble of the SM.
// A and B tiles are streamed global -> shared with 16-byte cp.async into a 4-stage ring
// (BK = 16 per stage).  Shared memory holds each operand in FRAGMENT-MAJOR order
//     [row-block of 8][k-pair of 8][lane 32] x double2
// so that one conflict-free LDS.128 per (row-block, k-pair) gives a thread the A (or B)
// fragment element of TWO consecutive DMMA k-steps: within a k-group of 8, step 0 contracts
// k = {0,2,4,6} and step 1 contracts k = {1,3,5,7}.  The same k permutation is applied to A
// and B, so the contraction is unchanged while every 16-byte global chunk (k, k+1 of one row)
// lands in exactly one double2 slot.
//
// On B200 every f64 mma.sync shape lowers to DMMA.8x8x4 (checked with cuobjdump), whose
// issue-rate peak measured 37.15 TFLOP/s (profiles/r01_ubench_fp64.txt); tcgen05.mma has no
// f64 kind, so this legacy-path instruction IS the FP64 tensor pipe on sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace be {

constexpr int BM = 128;
constexpr int BN = 64;
constexpr int BK = 16;
constexpr int KP = BK / 8;  // k-pairs per stage
constexpr int STAGES = 4;
constexpr int GEMM_THREADS = 128;
constexpr int GEMM_CTAS_PER_SM = 2;
constexpr int A_STAGE_D2 = BM * BK / 2;                  // double2 slots of the A operand per stage
constexpr int B_STAGE_D2 = BN * BK / 2;
constexpr int STAGE_D2 = A_STAGE_D2 + B_STAGE_D2;
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_D2 * 16;  // 98304

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;  // src-size 0 => zero fill, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Copies one operand stage (ROWS rows x 16 k) into fragment-major shared memory.
// g points at (row 0, k 0) of the tile; rows >= rows_valid are zero-filled.
template <int ROWS>
__device__ __forceinline__ void load_operand_stage(double2* sdst, const double* __restrict__ g, int ld,
                                                   int rows_valid, int tid) {
#pragma unroll
    for (int i = 0; i < (ROWS * BK / 2) / GEMM_THREADS; ++i) {
        int c = tid + i * GEMM_THREADS;
        int m = c >> 3;   // row in tile
        int kc = c & 7;   // 16-byte chunk along k
        bool v = m < rows_valid;
        const double* src = g + (size_t)(v ? m : 0) * ld + 2 * kc;
        int d = ((((m >> 3) * KP) + (kc >> 2)) << 5) + ((m & 7) << 2) + (kc & 3);
        cp_async16(sdst + d, src, v);
    }
}

struct TileAcc {
    double v[8][4][2];  // [m-block][n-block][pair]; warp tile 64 x 32
};

// Tried and dropped (r01g): interleaving the 8-row / 8-column blocks between the warps and skipping,
// under a per-warp 32-bit mask, the DMMAs of blocks past a ragged edge or strictly above the diagonal
// of a diagonal tile (~5% of the issued DMMAs at T=3012).  With both the masked and the plain loop
// inlined in one kernel ptxas hit the 255-register cap and spilled, and k_chol_update got 6% SLOWER
// (88.9 vs 83.5 ms per cfg2 step); the plain loop stays.
// Thread's coordinates inside the 128 x 128 tile for accumulator (mi, ni, e).
__device__ __forceinline__ int acc_row(int warp, int lane, int mi) { return (warp >> 1) * 64 + mi * 8 + (lane >> 2); }
__device__ __forceinline__ int acc_col(int warp, int lane, int ni) { return (warp & 1) * 32 + ni * 8 + 2 * (lane & 3); }

// acc (+)= sum over k in [0, klen) of A[i,k] * B[j,k] for a 128 (i) x 64 (j) tile; klen must be a multiple of BK (buffers are
// padded so that it always is).  A, B point at (tile row 0, k 0).  With ZERO_INIT = false the
// caller has preloaded acc (e.g. with -C, so that the epilogue is a pure store and the C read
// overlaps the pipeline prologue instead of serialising behind the mainloop).
template <bool ZERO_INIT = true>
__device__ __forceinline__ void gemm_nt_mainloop(const double* __restrict__ A, int lda, int a_rows,
                                                 const double* __restrict__ B, int ldb, int b_rows, int klen,
                                                 double2* smem, TileAcc& acc) {
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int wm = warp >> 1;
    const int wn = warp & 1;
    if (ZERO_INIT) {
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) acc.v[mi][ni][0] = acc.v[mi][ni][1] = 0.0;
    }

    const int ktiles = klen / BK;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < ktiles) {
            load_operand_stage<BM>(smem + s * STAGE_D2, A + s * BK, lda, a_rows, tid);
            load_operand_stage<BN>(smem + s * STAGE_D2 + A_STAGE_D2, B + s * BK, ldb, b_rows, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < ktiles; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < ktiles) {
                int s = nk % STAGES;
                load_operand_stage<BM>(smem + s * STAGE_D2, A + nk * BK, lda, a_rows, tid);
                load_operand_stage<BN>(smem + s * STAGE_D2 + A_STAGE_D2, B + nk * BK, ldb, b_rows, tid);
            }
            cp_async_commit();
        }
        const double2* sA = smem + (kt % STAGES) * STAGE_D2;
        const double2* sB = sA + A_STAGE_D2;
#pragma unroll
        for (int p = 0; p < KP; ++p) {
            double2 a[8], b[4];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) a[mi] = sA[((((wm * 8 + mi) * KP) + p) << 5) + lane];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = sB[((((wn * 4 + ni) * KP) + p) << 5) + lane];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc.v[mi][ni][0], acc.v[mi][ni][1], a[mi].x, b[ni].x);
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc.v[mi][ni][0], acc.v[mi][ni][1], a[mi].y, b[ni].y);
        }
    }
    cp_async_wait<0>();
    __syncthreads();  // every global read of A/B has landed: in-place epilogues are safe
}

// CTA rasterisation of the tile kernels: grid = tiles * 2 * B, TILE-major (the same tile of all B
// problems side by side, so every CTA of a wave has the same K length and the wave drains together).
// No operand is shared between co-resident CTAs in this order, which shows up as DRAM traffic well
// above the algorithmic bytes (k_lauum_cov: 133 GB per launch against 10.4 GB, profiles/r01f_traffic.csv)
// -- but DRAM then runs at ~2.8 TB/s, 43% of the measured peak, and is not the limiter: the
// problem-major order (-DBE_RASTER_PROBLEM_MAJOR), which lets the CTAs of one problem share operand
// rows through L2, measured 2% SLOWER on the cfg2 bench (30.08 vs 30.72 cells/s, same box, r01f).
#ifdef BE_RASTER_PROBLEM_MAJOR
__device__ __forceinline__ void cta_decode(int B, int& tile, int& half, int& b) {
    const int per = gridDim.x / B;
    b = blockIdx.x / per;
    const int r = blockIdx.x % per;
    tile = r >> 1;
    half = r & 1;
}
#else
__device__ __forceinline__ void cta_decode(int B, int& tile, int& half, int& b) {
    tile = blockIdx.x / (2 * B);
    const int rem = blockIdx.x % (2 * B);
    half = rem / B;
    b = rem % B;
}
#endif

// lower-triangular pair index -> (ti >= tj)
__device__ __forceinline__ void tri_decode(int idx, int& ti, int& tj) {
    int t = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
    while ((t + 1) * (t + 2) / 2 <= idx) ++t;
    while (t * (t + 1) / 2 > idx) --t;
    ti = t;
    tj = idx - t * (t + 1) / 2;
}

}  // namespace be
