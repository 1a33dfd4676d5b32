// fp64-EQUIVALENT GEMM on the INTEGER tensor cores of sm_100a (tcgen05.mma kind::i8, int32 accumulators in TMEM)
// by the Ozaki scheme -- the first kernel of DESIGN.md section 10 item 1, behind be_dgemm_nt_i8tc.  It is NOT on
// the fit -> weight -> barycentre path yet: the DMMA tile engine still carries the factorisation; this file is the
// tested building block the next round wires in.
//
// Scheme (S = 8 slices of 7 bits, exact in every step but the final fp64 sums):
//   row i of A:  a = 2^(E_i+1) x,  |x| < 1/2;   q_s = rint(128 x), x <- 128 x - q_s  (s = 1..S),  |q_s| <= 64 (int8)
//   A B^T = 2^(E_i+1) 2^(F_j+1) sum_{p=2}^{S+1} 128^-p I_p,     I_p = sum_{s+t=p} Q_s Q_t^T   (exact in int32)
// terms with s + t > S + 1 are dropped (relative size K 2^-55 of the row/column scale, the order of fp64 rounding).
// One CTA of six warps owns a 128 x 256 tile: warp 4 streams slice blocks with cp.async.bulk into an mbarrier ring
// (2/3/4/8 stages of 96/72/48/24 KB following the sweep), one thread of warp 5 issues the MMAs, warps 0-3 are the
// epilogue.  TMEM holds two 128 x 256 int32 accumulators, so the eight p are taken two at a time (p = 9,8 | 7,6 |
// 5,4 | 3,2): four sweeps over K; the two accumulators of a sweep are combined exactly in int64 (J = 128 I_lo + I_hi)
// and added to the fp64 partial sums of the tile (scratch laid out [tile][32-column block][row][32], L2-resident);
// the last sweep applies the power-of-two scales.  Slices live in global memory in the slab order
// [slice][tile][k/16][row][16 B], so that the 32-byte K block of a slice tile is contiguous and lands in shared
// memory directly in the canonical K-major no-swizzle UMMA layout.
// Measured stand-alone (tools/ozaki_dgemm.cu, profiles/r01p_ozaki_dgemm.txt): 66-74 TFLOP/s fp64-equivalent, error
// 6e-17 of sum|a||b| -- the DMMA kernels run at 28-30.
#pragma once
#include <stdint.h>

namespace be {
namespace oz {

constexpr int S = 8;  // slices
constexpr int TM = 128, TN = 256;
constexpr int A_SLICE_BYTES = 2 * TM * 16;  // one 32-byte K block of one slice: two 16-byte slabs x 128 rows
constexpr int B_SLICE_BYTES = 2 * TN * 16;
constexpr int SLICE_PAIR_BYTES = A_SLICE_BYTES + B_SLICE_BYTES;  // 12 KB: one slice of A and of B for a 32-byte K block
constexpr int RING_BYTES = 216 * 1024;
constexpr int MAXSTAGE = 8;
// stages of the ring in sweep g (ns = 8, 6, 4, 2 slices per operand): as many as fit -- 2, 3, 4, 8
__host__ __device__ constexpr int n_stages(int ns) { return RING_BYTES / (ns * SLICE_PAIR_BYTES) > MAXSTAGE ? MAXSTAGE : RING_BYTES / (ns * SLICE_PAIR_BYTES); }


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns of TMEM -> 32 registers per thread (no wait: see tcgen05.wait::ld)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}


// Asl [S][MT][K/16][128][16], Bsl [S][NT][K/16][256][16]; EA [M], FB [N] = exponents (scale 2^(E+1)); C [M][N] (zeroed)
__global__ void __launch_bounds__(192, 1)
k_ozaki_dgemm(const int8_t* __restrict__ Asl, const int8_t* __restrict__ Bsl, const int* __restrict__ EA,
              const int* __restrict__ FB, double* __restrict__ C, double* __restrict__ U, int M, int N, int K) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[MAXSTAGE];   // "the bulk copies of this stage have landed"
    __shared__ __align__(8) uint64_t bar_stage[MAXSTAGE];  // "the MMAs that read this stage are done"
    __shared__ __align__(8) uint64_t bar_acc;            // "both accumulators of this sweep are complete"
    __shared__ __align__(8) uint64_t bar_drained;        // "the 128 epilogue threads have read the accumulators"
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int MT = M / TM, NT = N / TN, nslab = K / 16, nkb = K / 32;
    const int mt = blockIdx.x % MT, nt = blockIdx.x / MT;
    if (tid == 0) {
        for (int s = 0; s < MAXSTAGE; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_stage[s])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_acc)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(smem_u32(&bar_drained)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = umma_idesc_i8(TM, TN);
    const uint32_t smem0 = smem_u32(smem);
    // (taking the sweeps in an order rotated per CTA, to spread the read-modify-write epilogues in time, measured 10 %
    // SLOWER: CTAs that share an operand tile stop reading the same slices at the same time and lose their L2 hits)
    constexpr int rot = 0;

    if (warp == 4) {
        // ===== PRODUCER (one warp): lane s copies slice s + 1 of A and of B; bytes are counted on bar_full =====
        uint32_t uses[MAXSTAGE] = {0, 0, 0, 0, 0, 0, 0, 0};  // times each stage has been filled (barrier phase = uses & 1)
        for (int g = 0; g < 4; ++g) {
            const int ns = S - 2 * ((g + rot) & 3), nst = n_stages(ns), stage_bytes = ns * SLICE_PAIR_BYTES;
            if (g > 0) mbar_wait(&bar_acc, (g - 1) & 1);  // the ring is re-partitioned: every MMA of the last sweep is done
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % nst;
                if (kb >= nst) mbar_wait(&bar_stage[st], (uses[st] - 1) & 1);  // the MMAs that read this stage are done
                const uint32_t sA = smem0 + st * stage_bytes, sB = sA + ns * A_SLICE_BYTES;
                const uint32_t fb = smem_u32(&bar_full[st]);
                if (lane == 0)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)stage_bytes) : "memory");
                __syncwarp();
                if (lane < ns) {
                    const int s = lane;
                    const int8_t* srcA = Asl + (((size_t)s * MT + mt) * nslab + 2 * kb) * (TM * 16);
                    const int8_t* srcB = Bsl + (((size_t)s * NT + nt) * nslab + 2 * kb) * (TN * 16);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     sA + s * A_SLICE_BYTES),
                                 "l"(srcA), "r"((uint32_t)A_SLICE_BYTES), "r"(fb)
                                 : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     sB + s * B_SLICE_BYTES),
                                 "l"(srcB), "r"((uint32_t)B_SLICE_BYTES), "r"(fb)
                                 : "memory");
                }
                uses[st] += 1;
            }
        }
    } else if (warp == 5) {
        // ===== MMA ISSUER (one thread) =====
        if (lane == 0) {
            uint32_t fulls[MAXSTAGE] = {0, 0, 0, 0, 0, 0, 0, 0};  // times each stage has been consumed
            for (int g = 0; g < 4; ++g) {
                const int p_hi = S + 1 - 2 * ((g + rot) & 3), p_lo = p_hi - 1, ns = p_hi - 1;
                const int nst = n_stages(ns), stage_bytes = ns * SLICE_PAIR_BYTES;
                if (g > 0) mbar_wait(&bar_drained, (g - 1) & 1);  // the epilogue has read the previous accumulators
                asm volatile("tcgen05.fence::after_thread_sync;");
                for (int kb = 0; kb < nkb; ++kb) {
                    const int st = kb % nst;
                    mbar_wait(&bar_full[st], fulls[st] & 1);
                    fulls[st] += 1;
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t sA = smem0 + st * stage_bytes, sB = sA + ns * A_SLICE_BYTES;
                    bool first_hi = kb == 0, first_lo = kb == 0;
                    for (int s = 1; s <= ns; ++s) {
                        const uint64_t ad = umma_desc(sA + (s - 1) * A_SLICE_BYTES, TM * 16, 128);
                        for (int pp = 0; pp < 2; ++pp) {
                            const int t = (pp == 0 ? p_hi : p_lo) - s;
                            if (t < 1 || t > ns) continue;
                            const uint64_t bd = umma_desc(sB + (t - 1) * B_SLICE_BYTES, TN * 16, 128);
                            bool& first = pp == 0 ? first_hi : first_lo;
                            const uint32_t accumulate = first ? 0u : 1u;
                            first = false;
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(
                                    tmem + (pp == 0 ? 0u : 256u)),
                                "l"(ad), "l"(bd), "r"(idesc), "r"(accumulate), "r"(0u));
                        }
                    }
                    umma_commit(&bar_stage[st]);
                    if (kb == nkb - 1) umma_commit(&bar_acc);
                }
            }
        }
    } else {
        // ===== EPILOGUE (warps 0-3, one TMEM lane quarter each): J = 128 I_lo + I_hi (exact), U += 2^(-7 p_hi) J =====
        for (int g = 0; g < 4; ++g) {
            const int p_hi = S + 1 - 2 * ((g + rot) & 3);
            mbar_wait(&bar_acc, g & 1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const int row = warp * 32 + lane;
            const int gi = mt * TM + row;
            double* crow = C + (size_t)gi * N + (size_t)nt * TN;
            const double w = ldexp(1.0, -7 * p_hi);
            const int ea = EA[gi];
            for (int c0 = 0; c0 < TN; c0 += 32) {
                uint32_t hi[32], lo[32];
                const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
                tmem_ld32(ta, hi);
                tmem_ld32(ta + 256u, lo);
                // the partial sums of the tile live in a scratch buffer laid out [tile][32-column block][row][32], so
                // that the 128 threads (= rows) of a step touch 32 KB of contiguous memory (the row-major C would put
                // them 8 N bytes apart); their loads are in flight together with the TMEM loads
                double* up = U + (((size_t)blockIdx.x * (TN / 32) + c0 / 32) * TM + row) * 32;
                double v[32];
                if (g > 0) {
#pragma unroll
                    for (int u = 0; u < 32; u += 4) {
                        const double4 q4 = *reinterpret_cast<const double4*>(up + u);
                        v[u] = q4.x; v[u + 1] = q4.y; v[u + 2] = q4.z; v[u + 3] = q4.w;
                    }
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c0 == TN - 32) {  // everything this thread needs has left TMEM: let the next sweep start
                    asm volatile("tcgen05.fence::before_thread_sync;");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_drained)) : "memory");
                }
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    const long long J = 128LL * (long long)(int32_t)lo[u] + (long long)(int32_t)hi[u];
                    v[u] = (g == 0 ? 0.0 : v[u]) + w * (double)J;
                }
                if (g < 3) {
#pragma unroll
                    for (int u = 0; u < 32; u += 4) *reinterpret_cast<double4*>(up + u) = make_double4(v[u], v[u + 1], v[u + 2], v[u + 3]);
                } else {
#pragma unroll
                    for (int u = 0; u < 32; u += 2) {
                        const int2 f2 = *reinterpret_cast<const int2*>(FB + nt * TN + c0 + u);
                        *reinterpret_cast<double2*>(crow + c0 + u) =
                            make_double2(ldexp(v[u], ea + f2.x + 2), ldexp(v[u + 1], ea + f2.y + 2));
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}



// Device-side slicing of a row-major fp64 matrix X [R][K] into S int8 slice arrays in slab order (tile height TR)
// and the row exponents.  One warp per row for the exponent (max |x| over the row), then one thread per
// (row, 16-column group): 128 bytes read, 16 bytes written per slice -- consecutive rows of a tile write
// consecutive 16-byte pieces of a slab.
__global__ void k_row_exponents(const double* __restrict__ X, int R, int K, int* __restrict__ E) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= R) return;
    double mx = 0.0;
    for (int k = lane; k < K; k += 32) mx = fmax(mx, fabs(X[(size_t)row * K + k]));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) E[row] = mx > 0.0 ? ilogb(mx) + 1 : 0;
}
__global__ void k_slice_rows(const double* __restrict__ X, int R, int K, int TR, const int* __restrict__ E,
                             int8_t* __restrict__ out) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nslab = K / 16;
    if (gid >= (size_t)R * nslab) return;
    // consecutive threads = consecutive rows of one slab: coalesced 16-byte stores
    const int slab = (int)((gid / TR) % nslab), rt = (int)(gid / ((size_t)TR * nslab)), r = (int)(gid % TR);
    const int row = rt * TR + r;
    const double* src = X + (size_t)row * K + slab * 16;
    const int e = E[row];
    double x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = ldexp(src[k], -(e + 1));
    const int RT = R / TR;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        union { int8_t b[16]; int4 v; } q;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const double y = x[k] * 128.0;
            const double qq = rint(y);
            x[k] = y - qq;
            q.b[k] = (int8_t)qq;
        }
        *reinterpret_cast<int4*>(out + ((((size_t)s * RT + rt) * nslab + slab) * TR + r) * 16) = q.v;
    }
}


}  // namespace oz
}  // namespace be
