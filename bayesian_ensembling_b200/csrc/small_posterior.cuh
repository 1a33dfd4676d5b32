// Small-T member kernels (T + 2 <= 256: BASELINE configs 1 and 4, the reference's own T = 86 / 165 fits).
//
// The blocked path of be_kernels.cuh needs ~17 launches per batch of member problems at T = 251, each a round trip
// through HBM, and its 128-wide diagonal-block kernel is the serial spine of every one of them (36 % of the cfg4
// step in round 1, tensor stage at 23 % of the DMMA peak).  Here ONE CTA owns ONE (cell, member) problem for a whole
// chain of stages, two CTAs per SM so that one CTA's latency-bound stretches (diagonal blocks, barriers, L2 round
// trips) are covered by the other CTA's tensor work:
//
//   k_small_factor_inverse :  M = C C^T (left-looking, 32-column block steps)  ->  u = C^-1 y (rides along as row T)
//                             ->  V = C^-T (upper, row-major)  ->  the posterior mean mu = y - E V u (kernel tail)
//   k_small_cov_factor     :  cov = D + E - E (V V^T) E  (lauum + the posterior epilogue of be_kernels.cuh)
//                             ->  scale_tri = chol(cov) with the rows (1, mu) riding along (data.py:38-39)
//
// (the Matern gram in front is k_matern32<1>, unchanged; the blocked path's k_posterior_mean pass over V is fused into
// the first kernel's tail.)
//
// Data stays in the per-problem workspace in global memory -- 2 x 512 KB per problem, L2-resident while the CTA works
// on it (296 CTAs x 1 MB is of the order of the 126 MB L2) -- and every stage is the same device routine: a "tall
// panel" product  OUT[rows, 32] (+)= A[rows, K] * B[32, K]^T  in which
//   * B (32 rows of the factor / of V, up to 256 long) is staged ONCE per block step in shared memory (cp.async),
//   * A is read by each warp straight from global memory as DMMA fragments (every lane loads 16 bytes: 8 rows x 64
//     contiguous bytes per instruction, all sectors fully used; each A element is used by exactly one warp, so a
//     shared-memory stage would add a barrier and nothing else), two k-groups ahead of the DMMAs that consume them,
//   * the accumulators (16 rows x 32 columns per warp pass) stay in registers through the triangular scale that
//     follows (X <- +/- OUT * Linv^T): with the k-permutation of dmma_gemm.cuh the accumulator pair of a thread IS
//     the A fragment of the next product, so no shuffle or shared-memory transpose is needed.
// The 32 x 32 diagonal blocks are factored and inverted in shared memory by the whole CTA (8-column steps: one warp
// factors the 8 x 8 pivot block redundantly per lane -- no shuffles -- and eliminates the rows, all threads update
// the trailing entries; the inverse is recursive doubling on the tensor pipe).  FP64 DMMA issues at one instruction per 16 cycles per SM sub-partition, so fragment
// traffic (one LDS.128 / LDG.128 per 2-4 DMMAs) is far from any limit: the design problem at this size is latency
// and barriers, not bandwidth.
//
// Layout: the padded [n, n] row-major buffers of be_kernels.cuh with n = ld = round_up(T + 2, 32); rows T, T+1 carry
// right-hand sides, the rest of the padding is identity; columns >= T never pivot.
#pragma once
#include "dmma_gemm.cuh"
#include "chol_diag.cuh"

namespace be {

constexpr int SB = 32;                  // block edge of the small-T blocked algorithms
constexpr int SM_MAX_DIM = 256;         // largest padded dimension the fused kernels take
constexpr int SM_LDB = SM_MAX_DIM + 8;  // B-panel row stride: == 8 (mod 16) doubles => conflict-free LDS.128 fragments
constexpr int SM_LDD = SB + 8;          // diagonal-block tiles, same residue
constexpr int SM_THREADS = 256;
constexpr int SM_WARPS = SM_THREADS / 32;
#ifndef BE_SMALL_DIAG_WARPS
#define BE_SMALL_DIAG_WARPS 2
#endif
constexpr int SM_DIAG_WARPS = BE_SMALL_DIAG_WARPS;
constexpr int SM_DIAG_THREADS = 32 * SM_DIAG_WARPS;
constexpr int SM_PROD_THREADS = SM_THREADS - SM_DIAG_THREADS;
constexpr int SM_LDT = 20;              // scratch tile of the recursive-doubling inverse (== 4 mod 16)
constexpr int SM_SMEM_DOUBLES = SB * SM_LDB + 3 * SB * SM_LDD + SB + 16 * SM_LDT + 2 + SM_MAX_DIM;  // + two mbarriers + yv
constexpr int SM_SMEM_BYTES = SM_SMEM_DOUBLES * 8;  // 102 912 B: two CTAs per SM

__host__ __device__ inline int small_dim(int T) { return ((T + 2 + SB - 1) / SB) * SB; }

// -DBE_SMALL_TIMING (tools/gpu_small_timing.sh builds a second library with it): per-phase clock64 totals of ONE CTA
// (block BE_SMALL_TIMING_BLOCK), read back through be_debug_small_timing.  Thread 0 speaks for the diagonal-block
// group, thread 128 for the product group.  slot = 16 * block step + 8 * group + phase.
#ifdef BE_SMALL_TIMING
#ifndef BE_SMALL_TIMING_BLOCK
#define BE_SMALL_TIMING_BLOCK 1000
#endif
__device__ long long g_small_timing[2][16 * 16];
__device__ __forceinline__ long long st_now() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    return t;
}
struct PhaseClock {
    long long t;
    int kern;
    __device__ explicit PhaseClock(int kernel_id) : t(st_now()), kern(kernel_id) {}
    __device__ void mark(int step, int phase) {
        if (blockIdx.x != BE_SMALL_TIMING_BLOCK || (threadIdx.x != 0 && threadIdx.x != 32 * BE_SMALL_DIAG_WARPS /* a thread of the product group (either mapping) */)) return;
        const long long now = st_now();
        g_small_timing[kern][16 * step + 8 * (threadIdx.x != 0) + phase] += now - t;
        t = now;
    }
};
#define ST_MARK(clk, step, phase) (clk).mark(step, phase)
#else
struct PhaseClock {
    __device__ explicit PhaseClock(int) {}
};
#define ST_MARK(clk, step, phase) ((void)0)
#endif

struct SmallSmem {
    double* B;    // [32][SM_LDB]  panel
    double* D;    // [32][SM_LDD]  diagonal block being factored
    double* Inv;  // [2][32][SM_LDD] inverses (lower) of diagonal blocks k (half k & 1) and k - 1
    double* rd;   // [32]          1 / diag
    double* Tmp;  // [16][SM_LDT]  scratch of the inverse
    unsigned long long* bars;  // [2] mbarriers of the panel loads: whole CTA, product group
    double* yv;   // [256]         the problem's y_var (kernel B's epilogue reads it per element)
    __device__ explicit SmallSmem(double* base)
        : B(base), D(base + SB * SM_LDB), Inv(base + SB * SM_LDB + SB * SM_LDD), rd(base + SB * SM_LDB + 3 * SB * SM_LDD),
          Tmp(base + SB * SM_LDB + 3 * SB * SM_LDD + SB),
          bars(reinterpret_cast<unsigned long long*>(base + SB * SM_LDB + 3 * SB * SM_LDD + SB + 16 * SM_LDT)),
          yv(base + SB * SM_LDB + 3 * SB * SM_LDD + SB + 16 * SM_LDT + 2) {}
};

// Phase parities of the two panel mbarriers, kept per thread (every thread that waits on a barrier flips its copy).
struct PanelPhase {
    unsigned parity[2];
};
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void panel_bars_init(const SmallSmem& sm, PanelPhase& ph) {
    ph.parity[0] = ph.parity[1] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(sm.bars)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(sm.bars + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
}

// 16 rows x 32 columns of output per warp pass, in DMMA accumulator layout: thread (g = lane / 4, q = lane % 4) holds
// rows 8 mi + g, columns 8 ni + 2 q and 8 ni + 2 q + 1.
struct SubAcc {
    double v[2][4][2];
};

__device__ __forceinline__ void sub_zero(SubAcc& a) {
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) a.v[mi][ni][0] = a.v[mi][ni][1] = 0.0;
}

// tile points at (row 0, column 0) of the 16 x 32 tile; works for global and shared memory alike
__device__ __forceinline__ void sub_load(SubAcc& a, const double* tile, int ld) {
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const double2 t = *reinterpret_cast<const double2*>(tile + (size_t)(8 * mi + g) * ld + 8 * ni + 2 * q);
            a.v[mi][ni][0] = t.x;
            a.v[mi][ni][1] = t.y;
        }
}

__device__ __forceinline__ void sub_store(const SubAcc& a, double* tile, int ld) {
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
            *reinterpret_cast<double2*>(tile + (size_t)(8 * mi + g) * ld + 8 * ni + 2 * q) =
                make_double2(a.v[mi][ni][0], a.v[mi][ni][1]);
}

// Asks for the 16 rows x 32 doubles at p (one 128-byte line per lane) without holding registers.  The products of
// these kernels are bound per warp -- a window lasts as long as its slowest warp -- and every 16 x 32 item begins with
// the L2 round trip of its first A fragments and (updates) ends with that of the tile it updates: prefetched into the
// L1 before the product, the latter leaves the warp's critical path (+1.3 % of the cfg4 step, r02Z).  Prefetching the
// NEXT item's leading A rows as well, the late update's tile and the scale phase's next tile measured 1.2 % slower
// than the update tile alone (r02a2): those stay plain loads.
__device__ __forceinline__ void prefetch_rows16(const double* p, int ld) {
    const int lane = threadIdx.x & 31;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p + (size_t)(lane >> 1) * ld + ((lane & 1) << 4)));
}

// acc += sum_{k in [k0, k1)} A[r, k] * B[c, k]   (r < 16, c < 32; k0, k1 multiples of 8)
// A: global, row-major, points at (tile row 0, column 0).  sB: shared, row c at sB + c * SM_LDB, its column kB0 at
// offset 0.  Within a k-group of 8 the first DMMA contracts k = {0,2,4,6}, the second {1,3,5,7} (dmma_gemm.cuh).
__device__ __forceinline__ void sub_gemm(SubAcc& acc, const double* A, int lda, const double* sB, int kB0, int k0, int k1) {
    if (k0 >= k1) return;
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const double* a0p = A + (size_t)g * lda + 2 * q;
    const double* a1p = a0p + (size_t)8 * lda;
    const double* bp = sB + g * SM_LDB + 2 * q - kB0;
    // A fragments run PF k-groups ahead of the DMMAs that consume them (L2 latency is ~800 cycles, a k-group is 16
    // DMMAs = 256 cycles of this sub-partition's FP64 pipe: two groups ahead left the pipe waiting, ncu r02e; three,
    // five and six groups ahead measured 0.8 / 1.6 / 2.3 % slower than four on the cfg4 step, r02g2)
    constexpr int PF = 4;
    double2 ra[PF][2];
#pragma unroll
    for (int p = 0; p < PF; ++p) {
        const int kk = min(k0 + 8 * p, k1 - 8);  // past the end: a harmless reload of the last group
        ra[p][0] = *reinterpret_cast<const double2*>(a0p + kk);
        ra[p][1] = *reinterpret_cast<const double2*>(a1p + kk);
    }
#pragma unroll 1
    for (int k = k0; k < k1; k += 8 * PF) {
#pragma unroll
        for (int p = 0; p < PF; ++p) {
            const int kc = k + 8 * p;
            if (kc < k1) {  // warp-uniform
                const double2 a0 = ra[p][0], a1 = ra[p][1];
                const int kn = min(kc + 8 * PF, k1 - 8);
                ra[p][0] = *reinterpret_cast<const double2*>(a0p + kn);
                ra[p][1] = *reinterpret_cast<const double2*>(a1p + kn);
                double2 b[4];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = *reinterpret_cast<const double2*>(bp + ni * 8 * SM_LDB + kc);
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    dmma884(acc.v[0][ni][0], acc.v[0][ni][1], a0.x, b[ni].x);
                    dmma884(acc.v[1][ni][0], acc.v[1][ni][1], a1.x, b[ni].x);
                }
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    dmma884(acc.v[0][ni][0], acc.v[0][ni][1], a0.y, b[ni].y);
                    dmma884(acc.v[1][ni][0], acc.v[1][ni][1], a1.y, b[ni].y);
                }
            }
        }
    }
}

// out = sign * t * Linv^T with Linv a LOWER-triangular 32 x 32 block in shared memory (row stride SM_LDD):
// out[r, c] = sign * sum_k t[r, k] Linv[c, k].  The accumulator pair (columns 8 nk + 2 q, + 1) of t is the A fragment
// of k-group nk; Linv[c, k] = 0 for k > c, so k-group nk only reaches the column blocks nj >= nk.
__device__ __forceinline__ void sub_scale(SubAcc& out, const SubAcc& t, const double* sInv, double sign) {
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    sub_zero(out);
#pragma unroll
    for (int nk = 0; nk < 4; ++nk)
#pragma unroll
        for (int nj = nk; nj < 4; ++nj) {
            const double2 b = *reinterpret_cast<const double2*>(sInv + (8 * nj + g) * SM_LDD + 8 * nk + 2 * q);
            dmma884(out.v[0][nj][0], out.v[0][nj][1], t.v[0][nk][0], b.x);
            dmma884(out.v[1][nj][0], out.v[1][nj][1], t.v[1][nk][0], b.x);
            dmma884(out.v[0][nj][0], out.v[0][nj][1], t.v[0][nk][1], b.y);
            dmma884(out.v[1][nj][0], out.v[1][nj][1], t.v[1][nk][1], b.y);
        }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            out.v[mi][ni][0] *= sign;
            out.v[mi][ni][1] *= sign;
        }
}

// 1 / sqrt(x) with a SHORT dependent chain: MUFU seed (relative error < 2^-22) and one third-order (Halley) step,
//   e = 1 - x y^2,  y <- y (1 + e / 2 + 3 e^2 / 8)      (error -> ~e^3 < 2^-65),
// four dependent FP64 operations instead of the six of two Newton steps (fast_rsqrt, chol_diag.cuh).  The pivot
// chain of the diagonal blocks is the serial spine of these kernels and every FP64 operation in it queues behind the
// DMMAs of the other warps on the same pipe.  x <= 0 gives NaN / inf like a failed LAPACK pivot would.
__device__ __forceinline__ double rsqrt_halley(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(e, 0.375, 0.5);
    const double ye = y * e;
    return fma(ye, p, y);
}

// A thread group inside the CTA: the whole CTA (barrier 0) or one of the two warp groups of the look-ahead windows
// (named barriers 1 and 2).  t / n = thread index in / size of the group, w / nw = warp index in / warps of it.
struct Grp {
    int t, n, w, nw, bar;
    __device__ __forceinline__ void sync() const {
        if (bar == 0)
            __syncthreads();
        else if (bar == 1)
            asm volatile("bar.sync 1, %0;" ::"n"(SM_DIAG_THREADS) : "memory");
        else
            asm volatile("bar.sync 2, %0;" ::"n"(SM_PROD_THREADS) : "memory");
    }
};
__device__ __forceinline__ Grp grp_cta() { return Grp{(int)threadIdx.x, SM_THREADS, (int)threadIdx.x >> 5, SM_WARPS, 0}; }
// Warp groups of the look-ahead windows: the diagonal-block group = the first SM_DIAG_WARPS warps (named barrier 1;
// the serial pivot chain lives in warp 0, the group's other phases are small), the product group = the rest (named
// barrier 2).  Measured at the cfg4 shape (profiles/r02*_bench_cfg4.json, per-phase tables r02f/g/h_small_timing.txt):
// 2 / 6: 12.8 k cells/s, 3 / 5: 11.7 k, 4 / 4: 12.1 k (tools/gpu_small_split.sh); a split that keeps every DMMA off
// warp 0's SM sub-partition (diagonal = warps 0, 1, 4: FP64 DMMA and DFMA share one pipe per sub-partition, and the
// pivot chain queues behind the DMMAs) made the pivot chain 30 % faster and the products slower: 11.7 k.
#ifdef BE_SMALL_DIAG_ON_SUBPARTITION0
// variant: diagonal group = warps 0 and 4 (both on SM sub-partition 0), product group = warps 1, 2, 3, 5, 6, 7: no
// DMMA is issued on sub-partition 0 during the windows (both CTAs of an SM use the same mapping)
__device__ __forceinline__ bool in_diag_group() { return (threadIdx.x & 127) < 32; }
__device__ __forceinline__ Grp grp_half() {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if ((w & 3) == 0) return Grp{(w >> 2) * 32 + lane, 64, w >> 2, 2, 1};
    const int gw = w < 4 ? w - 1 : w - 2;  // 1, 2, 3, 5, 6, 7 -> 0 .. 5
    return Grp{gw * 32 + lane, 192, gw, 6, 2};
}
#else
__device__ __forceinline__ bool in_diag_group() { return threadIdx.x < SM_DIAG_THREADS; }
__device__ __forceinline__ Grp grp_half() {
    const int tid = threadIdx.x;
    if (tid < SM_DIAG_THREADS) return Grp{tid, SM_DIAG_THREADS, tid >> 5, SM_DIAG_WARPS, 1};
    return Grp{tid - SM_DIAG_THREADS, SM_PROD_THREADS, (tid - SM_DIAG_THREADS) >> 5, SM_WARPS - SM_DIAG_WARPS, 2};
}
#endif

// 32 rows x klen columns (klen a multiple of 2) of a global row-major matrix -> the shared panel of group g, by TMA:
// one bulk copy (cp.async.bulk, no tensor map needed: every panel row is one contiguous run of klen * 8 bytes) per row,
// issued by the 32 lanes of the group's first warp and completed on the group's mbarrier, on which every thread of
// the group then waits.  The rows being read were written by this CTA's own (generic-proxy) stores, and the panel
// area itself is also written by ordinary stores between two loads: the issuing lanes order both against the async
// proxy with fence.proxy.async after the group barrier that precedes every call.
// (-DBE_SMALL_NO_TMA: the 16-byte cp.async loop of the first version, for A/B.)
__device__ __forceinline__ void load_panel(const Grp& g, double* sB, const double* src, int ld, int klen, const SmallSmem& sm,
                                           PanelPhase& ph) {
#ifdef BE_SMALL_NO_TMA
    const int cpr = klen >> 1;  // 16-byte chunks per row
    const int total = SB * cpr;
    for (int c = g.t; c < total; c += g.n) {
        const int r = c / cpr, kc = c - r * cpr;
        cp_async16(sB + r * SM_LDB + 2 * kc, src + (size_t)r * ld + 2 * kc, true);
    }
    cp_async_commit();
    cp_async_wait<0>();
#else
    const int which = g.bar == 0 ? 0 : 1;
    const unsigned bar = smem_addr(sm.bars + which);
    if (g.t < 32) {
        const unsigned bytes = (unsigned)klen * 8u;
        asm volatile("fence.proxy.async;" ::: "memory");
        if (g.t == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * SB) : "memory");
        __syncwarp();
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(sB + g.t * SM_LDB)),
                     "l"(src + (size_t)g.t * ld), "r"(bytes), "r"(bar)
                     : "memory");
    }
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(bar), "r"(ph.parity[which])
            : "memory");
    }
    ph.parity[which] ^= 1u;
#endif
}

// first sub-block (16 rows) >= lo that warp w of nw owns: sub-blocks are dealt round-robin
__device__ __forceinline__ int first_owned(int lo, int w, int nw) { return lo + (((w - lo) % nw + nw) % nw); }

// item t of a "snake" deal of count items to nw warps (even rounds ascending, odd rounds descending): with costs that
// fall linearly in the item index every warp's total is about the same.  Returns -1 past the end.
__device__ __forceinline__ int snake_item(int t, int nw, int count) {
    const int round = t / nw, pos = t - round * nw;
    const int r = (round & 1) ? round * nw + (nw - 1 - pos) : t;
    return r < count ? r : -1;
}

// Cholesky of the 32 x 32 block in sD (lower part, columns < nr real) by the threads of group g, in four 8-column steps:
//   phase 1 (first warp of the group, lane = row): every lane factors the 8 x 8 pivot block REDUNDANTLY in its
//            registers (no shuffles, no hand-off) and eliminates its own row with it -- the 8 dependent rsqrt of this
//            chain (~100 cycles each, profiles/r02c ubench: rsqrt 75, dfma 8.4) are the irreducible serial part;
//   phase 2 (all threads of the group): one trailing entry per thread, D[i, c] -= L[i, c0:c0+8] . L[c, c0:c0+8].
// Rows >= nr (padding / right-hand sides inside the band) are eliminated like any row below the real block; columns
// >= nr are never touched.  sRd receives 1 / diag (1 for the padding columns).  Returns the LAPACK-style report
// (0 = fine), valid in the first warp of the group.  Ends with a group barrier.
__device__ __forceinline__ int diag_factor32(const Grp& g, double* sD, double* sRd, int nr, int base, PhaseClock& clk, int step) {
    const int i = g.t & 31;
    int bad = 0;
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        const int c0 = 8 * s;
        const int w = min(8, nr - c0);
        if (w <= 0) {  // group-uniform
            if (g.t < 8) sRd[c0 + g.t] = 1.0;
            continue;
        }
        if (g.w == 0) {
            double Lb[8][8], pv[8], rs[8];
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b <= a; ++b) {
                    const double v = sD[(c0 + a) * SM_LDD + c0 + b];
                    Lb[a][b] = (a < w) ? v : (a == b ? 1.0 : 0.0);
                }
#pragma unroll
            for (int b = 0; b < 8; b += 2) {
                const double2 t = *reinterpret_cast<const double2*>(sD + i * SM_LDD + c0 + b);
                pv[b] = t.x;
                pv[b + 1] = t.y;
            }
            __syncwarp();  // every lane holds its copy of the pivot block before rows are rewritten
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double piv = Lb[j][j];
                bad = (bad == 0 && j < w && !(piv > 0.0)) ? base + c0 + j + 1 : bad;
                const double r = rsqrt_halley(piv);
                rs[j] = r;
#pragma unroll
                for (int a = j + 1; a < 8; ++a) Lb[a][j] *= r;
#pragma unroll
                for (int b = j + 1; b < 8; ++b)
#pragma unroll
                    for (int a = b; a < 8; ++a) Lb[a][b] = fma(-Lb[a][j], Lb[b][j], Lb[a][b]);
                pv[j] *= r;
#pragma unroll
                for (int b = j + 1; b < 8; ++b) pv[b] = fma(-pv[j], Lb[b][j], pv[b]);
            }
            if (i >= c0) {
#pragma unroll
                for (int b = 0; b < 8; b += 2) {
                    double2 o;
                    o.x = (c0 + b <= i) ? pv[b] : 0.0;
                    o.y = (c0 + b + 1 <= i) ? pv[b + 1] : 0.0;
                    *reinterpret_cast<double2*>(sD + i * SM_LDD + c0 + b) = o;
                }
            }
            if (i == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) sRd[c0 + j] = rs[j];
            }
        }
        ST_MARK(clk, step, 7);  // timing build only: pivot chain + row elimination (thread 0)
        g.sync();
        // trailing real columns [c0 + 8, nr): D[r, c] -= L[r, c0:c0+8] . L[c, c0:c0+8] for c <= r, one 8 x 8 tile per warp
        // pass on DMMA (two k-steps).  (One entry per thread with plain FMAs took 2.8 k cycles per step under the
        // other warps' DMMA traffic: profiles/r02p_small_timing.txt.)
        {
            const int lane = g.t & 31, gq = lane >> 2, q = lane & 3;
            const int p0 = c0 + 8, nt = (SB - p0) / 8;  // tile rows below the pivot block
            for (int task = g.w; task < nt * (nt + 1) / 2; task += g.nw) {
                int tr = 0;
                while ((tr + 1) * (tr + 2) / 2 <= task) ++tr;
                const int tc = task - tr * (tr + 1) / 2;
                const int r = p0 + 8 * tr + gq, cc = p0 + 8 * tc + 2 * q;
                const double* Pr = sD + (p0 + 8 * tr + gq) * SM_LDD + c0;
                const double* Pc = sD + (p0 + 8 * tc + gq) * SM_LDD + c0;
                double2 cv = *reinterpret_cast<const double2*>(sD + r * SM_LDD + cc);
                double acc0 = 0.0, acc1 = 0.0;
                dmma884(acc0, acc1, Pr[q], Pc[q]);
                dmma884(acc0, acc1, Pr[4 + q], Pc[4 + q]);
                cv.x = (cc < nr && cc <= r) ? cv.x - acc0 : cv.x;
                cv.y = (cc + 1 < nr && cc + 1 <= r) ? cv.y - acc1 : cv.y;
                *reinterpret_cast<double2*>(sD + r * SM_LDD + cc) = cv;
            }
        }
        g.sync();
        ST_MARK(clk, step, 6);  // timing build only: barrier + trailing entries + barrier (thread 0)
    }
    g.sync();
    return bad;
}

// OUT_p[s x s] = sign * X_p * Y_p for npairs pairs (row-major, "NN"), 8 x 8 output pieces dealt to the warps of g
__device__ __forceinline__ void pair_product(const Grp& g, const double* X, int ldx, int xs, const double* Y, int ldy, int ys,
                                             double* OUT, int ldo, int os, int s, int npairs, double sign) {
    const int lane = g.t & 31, gq = lane >> 2, q = lane & 3;
    const int fn = s / 8, per_pair = fn * fn;
    for (int task = g.w; task < npairs * per_pair; task += g.nw) {
        const int p = task / per_pair, rem = task - p * per_pair;
        const int fr = rem / fn, fc = rem - fr * fn;
        const double* A = X + (size_t)p * xs + fr * 8 * ldx;
        const double* Bm = Y + (size_t)p * ys + fc * 8;
        double acc[1][2] = {};
        frag_nn<1>(A, ldx, Bm, ldy, s, acc);
        *reinterpret_cast<double2*>(OUT + (size_t)p * os + (fr * 8 + gq) * ldo + fc * 8 + 2 * q) =
            make_double2(sign * acc[0][0], sign * acc[0][1]);
    }
}

// S = inverse of blockdiag(L11, I) (L11 = the nr real rows / columns of the factored block in sD), by the threads of
// group g: the four 8 x 8 diagonal blocks by forward substitution (one column per thread), then recursive doubling
//   inv([[A, 0], [B, C]]) = [[A^-1, 0], [-C^-1 B A^-1, C^-1]]
// at block sizes 8 and 16, every product on the FP64 tensor pipe.  Ends with a group barrier.
__device__ __forceinline__ void diag_inverse32(const Grp& g, const double* sD, const double* sRd, double* S, double* sTmp,
                                               int nr) {
    for (int e = g.t; e < SB * SB; e += g.n) {
        const int r = e >> 5, c = e & 31;
        S[r * SM_LDD + c] = c <= r ? (r < nr ? sD[r * SM_LDD + c] : (r == c ? 1.0 : 0.0)) : 0.0;
    }
    g.sync();
    {
        double x[8];
        const int blk = (g.t >> 3) & 3, cidx = g.t & 7;
        double* Lb = S + (blk * 8) * SM_LDD + blk * 8;
        if (g.t < 32) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double sacc = (i == cidx) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < i; ++k) sacc = fma(-Lb[i * SM_LDD + k], (k >= cidx) ? x[k] : 0.0, sacc);
                x[i] = (i >= cidx) ? sacc * sRd[blk * 8 + i] : 0.0;
            }
        }
        g.sync();
        if (g.t < 32) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i >= cidx) Lb[i * SM_LDD + cidx] = x[i];
        }
        g.sync();
    }
#pragma unroll 1
    for (int s = 8; s <= 16; s *= 2) {
        const int npairs = 16 / s;
        const int stride = 2 * s * SM_LDD + 2 * s;  // from one pair's A to the next
        // T_p = B_p * Ainv_p -> Tmp (pair p at column offset p * s, rows 0 .. s)
        pair_product(g, S + s * SM_LDD, SM_LDD, stride, S, SM_LDD, stride, sTmp, SM_LDT, s, s, npairs, 1.0);
        g.sync();
        // B_p = -Cinv_p * T_p
        pair_product(g, S + s * SM_LDD + s, SM_LDD, stride, sTmp, SM_LDT, s, S + s * SM_LDD, SM_LDD, stride, s, npairs, -1.0);
        g.sync();
    }
}

// One column step of the triangular inverse by the warps of group g, V = C^-T (upper, row-major):
//   V[j, i] = -(sum_{p=j}^{i-1} V[j, p] C[i, p]^T) Linv_i^T,  j < i   (the diagonal tile V[i, i] is already in place).
// Loads the panel C[i, 0:32 i) into sm.B itself; sInv = Linv_i.  Begins and ends with a group barrier on sm.B.
__device__ __forceinline__ void trtri_column(const Grp& g, double* Vt, const double* Cm, int ld, int i, const double* sInv,
                                             const SmallSmem& sm, PanelPhase& ph) {
    const int ic = SB * i;
    load_panel(g, sm.B, Cm + (size_t)ic * ld, ld, ic, sm, ph);
    g.sync();
    // row sub-block r contracts K = 32 (i - r / 2) columns: snake deal, so that every warp's total K is about the same
#pragma unroll 1
    for (int t = g.w; t < 2 * i + g.nw; t += g.nw) {
        const int r = snake_item(t, g.nw, 2 * i);
        if (r < 0) continue;
        SubAcc acc, x;
        sub_zero(acc);
        sub_gemm(acc, Vt + (size_t)16 * r * ld, ld, sm.B, 0, SB * (r >> 1), ic);
        sub_scale(x, acc, sInv, -1.0);
        sub_store(x, Vt + (size_t)16 * r * ld + ic, ld);
    }
    g.sync();
}

// Early part of the left-looking update of block column kn ("part A": the columns [0, kend) of the factor that are
// final while the diagonal block kn - 1 is still being factored), by the warps of group g:
//   Mat[r, kn] -= Mat[r, 0:kend) Mat[kn, 0:kend)^T   for every 16-row sub-block r of block rows >= kn, in place.
__device__ __forceinline__ void update_early(const Grp& g, double* Mat, int ld, int nb, int kn, int kend, const SmallSmem& sm,
                                             PanelPhase& ph) {
    load_panel(g, sm.B, Mat + (size_t)SB * kn * ld, ld, kend, sm, ph);
    g.sync();
#pragma unroll 1
    for (int r = first_owned(2 * kn, g.w, g.nw); r < 2 * nb; r += g.nw) {
        SubAcc acc, c;
        sub_zero(acc);
        double* tile = Mat + (size_t)16 * r * ld + SB * kn;
        // the tile being updated is needed after the product: asked for now (loading it into registers BEFORE the
        // product cost 100 bytes of spills and 2 %: r02Y)
        prefetch_rows16(tile, ld);
        sub_gemm(acc, Mat + (size_t)16 * r * ld, ld, sm.B, 0, 0, kend);
        sub_load(c, tile, ld);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                c.v[mi][ni][0] -= acc.v[mi][ni][0];
                c.v[mi][ni][1] -= acc.v[mi][ni][1];
            }
        sub_store(c, tile, ld);
    }
}

// One block column of lauum with the posterior epilogue (semantics of k_lauum_cov, be_kernels.cuh) by the warps of g:
//   cov[i, j] = D + E - E (sum_{k >= 32 i} V[i, k] V[j, k]) E  for the sub-blocks of block rows i >= j; Work receives the
// padded lower covariance, rows T / T+1 = (1, mu), identity padding.  Loads its panel V[j, 32 j : n) into sm.B.
struct CovEpilogue {
    const double* yv;   // y_var of this problem [T]
    const double* mu;   // posterior mean [T]
    double* var_diag;   // [T]
    double* cov_dense;  // [T][T] or nullptr
    double jitter;
    int T;
};
__device__ __forceinline__ void lauum_column(const Grp& g, const double* Vb, double* Wb, int n, int nb, int j,
                                             const CovEpilogue& ep, const SmallSmem& sm, PanelPhase& ph) {
    const int lane = g.t & 31, gq = lane >> 2, q = lane & 3;
    const int jc = SB * j, T = ep.T;
    load_panel(g, sm.B, Vb + (size_t)jc * n + jc, n, n - jc, sm, ph);
    g.sync();
#pragma unroll 1
    for (int t = g.w; t < 2 * (nb - j) + g.nw; t += g.nw) {
        // sub-block 2 j + it contracts K = n - 32 (j + it / 2) columns: snake deal again
        const int it = snake_item(t, g.nw, 2 * (nb - j));
        if (it < 0) continue;
        const int r = 2 * j + it;
        SubAcc acc;
        sub_zero(acc);
        sub_gemm(acc, Vb + (size_t)16 * r * n, n, sm.B, jc, SB * (r >> 1), n);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
            const int gr = 16 * r + 8 * mi + gq;
            const double dr = gr < T ? ep.yv[gr] : 0.0;
            const double er = dr + ep.jitter;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                double out[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int gc = jc + 8 * ni + 2 * q + e;
                    double val;
                    if (gr < T && gc < T) {
                        const double ec = ep.yv[gc] + ep.jitter;
                        val = -er * ec * acc.v[mi][ni][e];
                        if (gr == gc) {
                            val += dr + er;
                            ep.var_diag[gr] = val;
                        }
                        if (ep.cov_dense && gc <= gr) {
                            ep.cov_dense[(size_t)gr * T + gc] = val;
                            ep.cov_dense[(size_t)gc * T + gr] = val;
                        }
                    } else if (gr == T && gc < T) {
                        val = 1.0;
                    } else if (gr == T + 1 && gc < T) {
                        val = ep.mu[gc];
                    } else {
                        val = gr == gc ? 1.0 : 0.0;
                    }
                    out[e] = val;
                }
                *reinterpret_cast<double2*>(Wb + (size_t)gr * n + jc + 8 * ni + 2 * q) = make_double2(out[0], out[1]);
            }
        }
    }
    g.sync();
}

// In-place Cholesky of the lower triangle of the padded n x n matrix Mat (n = 32 nb), real dimension T, left-looking in
// 32-column block steps WITH LOOK-AHEAD.  The update of block column k + 1 is split in two: its early part (all
// columns of the factor but the last 32) does not depend on the diagonal block k, so it runs -- on warps 4-7, one per
// SM sub-partition -- in the same "window" in which warps 0-3 factor and invert diagonal block k (the serial spine:
// 32 dependent rsqrt per block).  The window also takes independent work of the caller's (`extra`: the triangular
// inverse's column k - 1 in kernel A, lauum's column k + 1 in kernel B).  After the window the whole CTA scales
// column k by the inverse (panel TRSM recast as a product) and applies the late part (K = 32) of the update of column
// k + 1, which leaves the next diagonal block in shared memory.  Three CTA barriers per step.
// If Vt != nullptr the diagonal tiles of V = C^-T are written as well.
// On return sm.Inv (ping-pong half (nb - 1) & 1) still holds the inverse of the last diagonal block.
template <class Extra>
__device__ __forceinline__ void potrf_small(double* Mat, int ld, int nb, int T, double* Vt, int* info_b, const SmallSmem& sm,
                                            Extra extra, PhaseClock& clk, PanelPhase& ph) {
    const Grp cta = grp_cta(), half = grp_half();
    const bool diag_group = in_diag_group();
    {   // diagonal block 0
        SubAcc c;
        if (cta.w < 2) {
            sub_load(c, Mat + (size_t)16 * cta.w * ld, ld);
            sub_store(c, sm.D + cta.w * 16 * SM_LDD, SM_LDD);
        }
    }
    __syncthreads();
    ST_MARK(clk, 15, 0);
#pragma unroll 1
    for (int k = 0; k < nb; ++k) {
        const int kc = SB * k;
        const int nr = max(0, min(SB, T - kc));
        double* sInv = sm.Inv + (k & 1) * SB * SM_LDD;
        // ---- window: diagonal block k  ||  independent products
        if (diag_group) {
            const int bad = diag_factor32(half, sm.D, sm.rd, nr, kc, clk, k);
            ST_MARK(clk, k, 6);
            if (threadIdx.x == 0 && bad != 0 && info_b && *info_b == 0) *info_b = bad;
            diag_inverse32(half, sm.D, sm.rd, sInv, sm.Tmp, nr);
        } else {
            extra(half, k);
            ST_MARK(clk, k, 6);
            if (k >= 1 && k + 1 < nb) update_early(half, Mat, ld, nb, k + 1, kc, sm, ph);
        }
        ST_MARK(clk, k, 0);  // the group's own work in the window
        __syncthreads();
        ST_MARK(clk, k, 1);  // waiting for the other group
        // ---- the factor's diagonal block (lower, real columns only; strict upper part of the real rows cleaned) and
        // the diagonal tile of V = C^-T
        for (int e = threadIdx.x; e < SB * SB; e += SM_THREADS) {
            const int r = e >> 5, c = e & 31;
            double* dst = Mat + (size_t)(kc + r) * ld + kc + c;
            if (c < nr && c <= r)
                *dst = sm.D[r * SM_LDD + c];
            else if (r < nr && c > r)
                *dst = 0.0;
            if (Vt) Vt[(size_t)(kc + r) * ld + kc + c] = r <= c ? sInv[c * SM_LDD + r] : 0.0;
        }
        // ---- panel below the diagonal block: L[r, k] = unscaled * Linv^T; the rows of block k + 1 also go to sm.B,
        // where they are the B operand of the late update
#pragma unroll 1
        for (int r = first_owned(2 * k + 2, cta.w, cta.nw); r < 2 * nb; r += cta.nw) {
            SubAcc t, x;
            sub_load(t, Mat + (size_t)16 * r * ld + kc, ld);
            sub_scale(x, t, sInv, 1.0);
            sub_store(x, Mat + (size_t)16 * r * ld + kc, ld);
            if (r < 2 * k + 4) sub_store(x, sm.B + (r - 2 * k - 2) * 16 * SM_LDB, SM_LDB);
        }
        ST_MARK(clk, k, 2);  // diagonal tile + panel scale
        __syncthreads();
        ST_MARK(clk, k, 3);
        // ---- late part of the update of column k + 1 (K = the 32 columns just scaled)
        if (k + 1 < nb) {
#pragma unroll 1
            for (int r = first_owned(2 * k + 2, cta.w, cta.nw); r < 2 * nb; r += cta.nw) {
                SubAcc acc, c;
                sub_zero(acc);
                sub_gemm(acc, Mat + (size_t)16 * r * ld, ld, sm.B, kc, kc, kc + SB);  // A = the rows this warp just stored
                double* tile = Mat + (size_t)16 * r * ld + kc + SB;
                sub_load(c, tile, ld);
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) {
                        c.v[mi][ni][0] -= acc.v[mi][ni][0];
                        c.v[mi][ni][1] -= acc.v[mi][ni][1];
                    }
                if (r < 2 * k + 4)
                    sub_store(c, sm.D + (r - 2 * k - 2) * 16 * SM_LDD, SM_LDD);
                else
                    sub_store(c, tile, ld);
            }
        }
        ST_MARK(clk, k, 4);  // late update
        __syncthreads();
        ST_MARK(clk, k, 5);
    }
}

// mu_i = y_i - E_i sum_{k >= i} V[i, k] u[k]  (V = C^-T upper, rows k-contiguous; E = y_var + jitter).  One warp per row,
// FOUR rows of a warp in flight at a time (the rows are short -- T - i elements -- so one row at a time is a chain of
// L2 round trips and five shuffles: ~40 k cycles per CTA).  u is staged in the panel buffer,
// free at this point: every bulk copy into it was waited for by the product that used it.  Not inlined: the kernel body
// sits at the 128-register cap.
__device__ __noinline__ void small_posterior_mean(const double* __restrict__ Vb, int n, int T, const double* __restrict__ u,
                                                  const double* __restrict__ y_mean, const double* __restrict__ y_var,
                                                  double jitter, double* __restrict__ mu, double* us) {
    for (int j = threadIdx.x; j < T; j += SM_THREADS) us[j] = u[j];
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = SM_THREADS / 32;
    for (int i0 = w; i0 < T; i0 += 4 * NW) {
        // all sixteen 16-byte loads of the four rows are issued before the first use (T <= 254: at most four chunks of
        // 64 per row); written as loops over k they would be four dependent L2 round trips per row, one row after another
        double2 v[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = i0 + r * NW;
            const double* row = Vb + (size_t)min(i, T - 1) * n;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k = (i & ~1) + 2 * lane + 64 * c;  // aligned start; element k < i (if any) is an explicit zero
                v[r][c] = (i < T && k < T) ? *reinterpret_cast<const double2*>(row + k) : make_double2(0.0, 0.0);
            }
        }
        double sum[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = i0 + r * NW;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k = (i & ~1) + 2 * lane + 64 * c;
                if (i < T && k < T) {
                    if (k >= i) s0 += v[r][c].x * us[k];
                    if (k + 1 < T && k + 1 >= i) s1 += v[r][c].y * us[k + 1];
                }
            }
            sum[r] = s0 + s1;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int r = 0; r < 4; ++r) sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], o);
        }
        if (lane < 4) {
            const int i = i0 + lane * NW;
            if (i < T) {
                const double s = lane == 0 ? sum[0] : lane == 1 ? sum[1] : lane == 2 ? sum[2] : sum[3];
                mu[i] = y_mean[i] - (y_var[i] + jitter) * s;
            }
        }
    }
}

// ---- kernel A: M = C C^T, u = C^-1 y (row T), V = C^-T ------------------------------------------------------
// Mat [B][n][n] holds the lower tiles of M = K + diag(y_var + jitter) with row T = y_mean (k_matern32<1>); on return
// it holds C (row T zeroed), Vt holds C^-T, u [B][T] the forward-substituted right-hand side.  The triangular
// inverse's column k - 1 runs in the window of diagonal block k; its last column after the loop.
__global__ void __launch_bounds__(SM_THREADS, 2)
    k_small_factor_inverse(double* Mat, double* Vt, double* u, int* info, int n, int T,
                           const double* __restrict__ y_mean, const double* __restrict__ y_var, double jitter,
                           double* __restrict__ mu) {
    extern __shared__ __align__(16) double small_smem[];
    const SmallSmem sm(small_smem);
    const int b = blockIdx.x, nb = n / SB;
    double* Mb = Mat + (size_t)b * n * n;
    double* Vb = Vt + (size_t)b * n * n;
    // Row T (the right-hand side y, becoming u) must be zeroed before a column of the triangular inverse reads the block
    // row it lives in: block tb = T / 32 -- the last block, or the one before it when T + 1 is a multiple of 32.
    const int tb = T / SB;
    PhaseClock clk(0);
    PanelPhase ph;
    panel_bars_init(sm, ph);
    potrf_small(Mb, n, nb, T, Vb, info + b, sm, [&](const Grp& g, int k) {
        if (k >= 2 && k - 1 < tb) trtri_column(g, Vb, Mb, n, k - 1, sm.Inv + ((k - 1) & 1) * SB * SM_LDD, sm, ph);
    }, clk, ph);
    for (int j = threadIdx.x; j < T; j += SM_THREADS) {
        double* p = Mb + (size_t)T * n + j;
        u[(size_t)b * T + j] = *p;
        *p = 0.0;  // the triangular inverse sees blockdiag(C, I)
    }
    __syncthreads();
    ST_MARK(clk, 15, 1);
    // the remaining columns (normally just the last one).  Only the inverses of the last two diagonal blocks are still
    // in shared memory, which is all that can be asked for: tb >= nb - 2.
    for (int i = max(1, tb); i < nb; ++i) trtri_column(grp_cta(), Vb, Mb, n, i, sm.Inv + (i & 1) * SB * SM_LDD, sm, ph);
    ST_MARK(clk, 15, 2);
    // ---- posterior mean  mu = y - E (V u): the arithmetic (and its order) of k_posterior_mean, which cost a launch and a
    // pass over V of its own (0.85 ms = 4.4 % of the cfg4 step).  Here it costs ~0.55 ms of this kernel (r02D/E: its ~60
    // FP64 additions and products per warp and row group queue behind the other CTA's DMMAs like everything else on
    // the shared FP64 pipe; an L2 prefetch of V under the last column changed nothing, so it is not the memory).
    __syncthreads();
    small_posterior_mean(Vb, n, T, u + (size_t)b * T, y_mean + (size_t)b * T, y_var + (size_t)b * T, jitter,
                         mu + (size_t)b * T, sm.B);
    ST_MARK(clk, 15, 3);
}

// ---- kernel B: cov = D + E - E (V V^T) E, then scale_tri = chol(cov) with (1, mu) riding along ---------------------
// lauum's block column 0 is formed by the whole CTA, column k + 1 in the window of diagonal block k (it only reads V).
__global__ void __launch_bounds__(SM_THREADS, 2)
    k_small_cov_factor(const double* Vt, double* Work, const double* __restrict__ y_var, double jitter,
                       const double* __restrict__ mu, double* __restrict__ var_diag, double* __restrict__ cov_dense,
                       int* info, int n, int T) {
    extern __shared__ __align__(16) double small_smem[];
    const SmallSmem sm(small_smem);
    const int b = blockIdx.x, nb = n / SB;
    const double* Vb = Vt + (size_t)b * n * n;
    double* Wb = Work + (size_t)b * n * n;
    // the epilogue of every 16 x 32 item reads y_var of its 2 rows and 8 column pairs: from shared memory, not as
    // global loads whose latency sits at the end of each item of each warp (a window lasts as long as its slowest
    // warp): 8.74 -> 8.35 ms per 10 240 problems (r02X)
    for (int j = threadIdx.x; j < T; j += SM_THREADS) sm.yv[j] = y_var[(size_t)b * T + j];
    const CovEpilogue ep{sm.yv, mu + (size_t)b * T, var_diag + (size_t)b * T,
                         cov_dense ? cov_dense + (size_t)b * T * T : nullptr, jitter, T};
    PhaseClock clk(1);
    PanelPhase ph;
    panel_bars_init(sm, ph);
    lauum_column(grp_cta(), Vb, Wb, n, nb, 0, ep, sm, ph);
    ST_MARK(clk, 15, 3);
    potrf_small(Wb, n, nb, T, nullptr, info + b, sm, [&](const Grp& g, int k) {
        if (k + 1 < nb) lauum_column(g, Vb, Wb, n, nb, k + 1, ep, sm, ph);
    }, clk, ph);
}

}  // namespace be
