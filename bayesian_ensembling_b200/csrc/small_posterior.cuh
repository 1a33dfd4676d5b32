// Small-T member kernels (T + 2 <= 256: BASELINE configs 1 and 4, the reference's own T = 86 / 165 fits).
//
// The blocked path of be_kernels.cuh needs ~17 launches per batch of member problems at T = 251, each a round trip
// through HBM, and its 128-wide diagonal-block kernel is the serial spine of every one of them (36 % of the cfg4
// step in round 1, tensor stage at 23 % of the DMMA peak).  Here ONE CTA owns ONE (cell, member) problem for a whole
// chain of stages, two CTAs per SM so that one CTA's latency-bound stretches (diagonal blocks, barriers, L2 round
// trips) are covered by the other CTA's tensor work:
//
//   k_small_factor_inverse :  M = C C^T (left-looking, 32-column block steps)  ->  u = C^-1 y (rides along as row T)
//                             ->  V = C^-T (upper, row-major)
//   k_small_cov_factor     :  cov = D + E - E (V V^T) E  (lauum + the posterior epilogue of be_kernels.cuh)
//                             ->  scale_tri = chol(cov) with the rows (1, mu) riding along (data.py:38-39)
//
// (the posterior mean between the two is k_posterior_mean, the Matern gram in front is k_matern32<1>: both unchanged).
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
constexpr int SM_LDT = 20;              // scratch tile of the recursive-doubling inverse (== 4 mod 16)
constexpr int SM_SMEM_DOUBLES = SB * SM_LDB + 2 * SB * SM_LDD + SB + 16 * SM_LDT;
constexpr int SM_SMEM_BYTES = SM_SMEM_DOUBLES * 8;  // 90 624 B: two CTAs per SM

__host__ __device__ inline int small_dim(int T) { return ((T + 2 + SB - 1) / SB) * SB; }

struct SmallSmem {
    double* B;    // [32][SM_LDB]  panel
    double* D;    // [32][SM_LDD]  diagonal block being factored
    double* Inv;  // [32][SM_LDD]  its inverse (lower)
    double* rd;   // [32]          1 / diag
    double* Tmp;  // [16][SM_LDT]  scratch of the inverse
    __device__ explicit SmallSmem(double* base)
        : B(base), D(base + SB * SM_LDB), Inv(base + SB * SM_LDB + SB * SM_LDD), rd(base + SB * SM_LDB + 2 * SB * SM_LDD),
          Tmp(base + SB * SM_LDB + 2 * SB * SM_LDD + SB) {}
};

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

// acc += sum_{k in [k0, k1)} A[r, k] * B[c, k]   (r < 16, c < 32; k0, k1 multiples of 8)
// A: global, row-major, points at (tile row 0, column 0).  sB: shared, row c at sB + c * SM_LDB, its column kB0 at
// offset 0.  Within a k-group of 8 the first DMMA contracts k = {0,2,4,6}, the second {1,3,5,7} (dmma_gemm.cuh).
__device__ __forceinline__ void sub_gemm(SubAcc& acc, const double* A, int lda, const double* sB, int kB0, int k0, int k1) {
    if (k0 >= k1) return;
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const double* a0p = A + (size_t)g * lda + 2 * q;
    const double* a1p = a0p + (size_t)8 * lda;
    const double* bp = sB + g * SM_LDB + 2 * q - kB0;
    // two k-groups of A in flight ahead of the one being multiplied
    double2 a0 = *reinterpret_cast<const double2*>(a0p + k0);
    double2 a1 = *reinterpret_cast<const double2*>(a1p + k0);
    double2 n0 = a0, n1 = a1;
    if (k0 + 8 < k1) {
        n0 = *reinterpret_cast<const double2*>(a0p + k0 + 8);
        n1 = *reinterpret_cast<const double2*>(a1p + k0 + 8);
    }
#pragma unroll 2
    for (int k = k0; k < k1; k += 8) {
        double2 m0 = n0, m1 = n1;
        if (k + 16 < k1) {
            m0 = *reinterpret_cast<const double2*>(a0p + k + 16);
            m1 = *reinterpret_cast<const double2*>(a1p + k + 16);
        }
        double2 b[4];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) b[ni] = *reinterpret_cast<const double2*>(bp + ni * 8 * SM_LDB + k);
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
        a0 = n0;
        a1 = n1;
        n0 = m0;
        n1 = m1;
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

// 32 rows x klen columns (klen a multiple of 2) of a global row-major matrix -> the shared panel; the caller syncs
__device__ __forceinline__ void load_panel(double* sB, const double* src, int ld, int klen) {
    const int cpr = klen >> 1;  // 16-byte chunks per row
    const int total = SB * cpr;
    for (int c = threadIdx.x; c < total; c += SM_THREADS) {
        const int r = c / cpr, kc = c - r * cpr;
        cp_async16(sB + r * SM_LDB + 2 * kc, src + (size_t)r * ld + 2 * kc, true);
    }
    cp_async_commit();
    cp_async_wait<0>();
}

// first sub-block (16 rows) >= lo that this warp owns: sub-blocks are dealt to the 8 warps round-robin, so that the
// shrinking (potrf, lauum) or growing (trtri) set of active rows of a block step is spread over all warps
__device__ __forceinline__ int first_owned(int lo, int warp) { return lo + ((warp - lo) & (SM_WARPS - 1)); }

// Cholesky of the 32 x 32 block in sD (lower part, columns < nr real) by the whole CTA, in four 8-column steps:
//   phase 1 (warp 0, lane = row): every lane factors the 8 x 8 pivot block REDUNDANTLY in its registers (no
//            shuffles, no hand-off) and eliminates its own row with it -- the 8 dependent rsqrt of this chain
//            (~100 cycles each, profiles/r02c ubench: rsqrt 75, dfma 8.4) are the irreducible serial part;
//   phase 2 (all 256 threads): one trailing entry per thread, D[i, c] -= L[i, c0:c0+8] . L[c, c0:c0+8].
// (The first version did phase 2 inside warp 0, one dependent chain per column: 15 k cycles per block, with seven
// warps waiting at the barrier -- 38 % of the kernel, ncu r02c.)  Rows >= nr (padding / right-hand sides inside the
// band) are eliminated like any row below the real block; columns >= nr are never touched.  sRd receives 1 / diag
// (1 for the padding columns).  Returns the LAPACK-style report (0 = fine), valid in warp 0.  Ends with a barrier.
__device__ __forceinline__ int diag_factor32(double* sD, double* sRd, int nr, int base) {
    const int tid = threadIdx.x, i = tid & 31, warp = tid >> 5;
    int bad = 0;
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        const int c0 = 8 * s;
        const int w = min(8, nr - c0);
        if (w <= 0) {  // CTA-uniform
            if (tid < 8) sRd[c0 + tid] = 1.0;
            continue;
        }
        if (warp == 0) {
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
                const double r = fast_rsqrt(piv);
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
        __syncthreads();
        // trailing real columns [c0 + 8, nr): one entry (row, col <= row) per thread
        const int m = SB - c0 - 8;
        for (int e = tid; e < m * m; e += SM_THREADS) {
            const int ii = e / m, cc = e - ii * m;
            const int r = c0 + 8 + ii, c = c0 + 8 + cc;
            if (c > r || c >= nr) continue;
            const double* lr = sD + r * SM_LDD + c0;
            const double* lc = sD + c * SM_LDD + c0;
            double acc0 = sD[r * SM_LDD + c], acc1 = 0.0;
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
                const double2 a0 = *reinterpret_cast<const double2*>(lr + j), a1 = *reinterpret_cast<const double2*>(lr + j + 2);
                const double2 b0 = *reinterpret_cast<const double2*>(lc + j), b1 = *reinterpret_cast<const double2*>(lc + j + 2);
                acc0 = fma(-a0.x, b0.x, acc0);
                acc1 = fma(-a0.y, b0.y, acc1);
                acc0 = fma(-a1.x, b1.x, acc0);
                acc1 = fma(-a1.y, b1.y, acc1);
            }
            sD[r * SM_LDD + c] = acc0 + acc1;
        }
        __syncthreads();
    }
    __syncthreads();
    return bad;
}

// sInv = inverse of blockdiag(L11, I) (L11 = the nr real rows / columns of the factored block in sD), by the whole
// CTA: the four 8 x 8 diagonal blocks by forward substitution (one column per thread), then recursive doubling
//   inv([[A, 0], [B, C]]) = [[A^-1, 0], [-C^-1 B A^-1, C^-1]]
// at block sizes 8 and 16, every product on the FP64 tensor pipe (level_product of chol_diag.cuh).  The first
// version solved one column per lane in ONE warp: a 496-FMA dependent chain, 10 k cycles per block (ncu r02c).
// Ends with a barrier.
__device__ __forceinline__ void diag_inverse32(const double* sD, const double* sRd, double* S, double* sTmp, int nr) {
    const int tid = threadIdx.x;
    for (int e = tid; e < SB * SB; e += SM_THREADS) {
        const int r = e >> 5, c = e & 31;
        S[r * SM_LDD + c] = c <= r ? (r < nr ? sD[r * SM_LDD + c] : (r == c ? 1.0 : 0.0)) : 0.0;
    }
    __syncthreads();
    {
        double x[8];
        const int blk = (tid >> 3) & 3, cidx = tid & 7;
        double* Lb = S + (blk * 8) * SM_LDD + blk * 8;
        if (tid < 32) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double sacc = (i == cidx) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < i; ++k) sacc = fma(-Lb[i * SM_LDD + k], (k >= cidx) ? x[k] : 0.0, sacc);
                x[i] = (i >= cidx) ? sacc * sRd[blk * 8 + i] : 0.0;
            }
        }
        __syncthreads();
        if (tid < 32) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i >= cidx) Lb[i * SM_LDD + cidx] = x[i];
        }
        __syncthreads();
    }
#pragma unroll 1
    for (int s = 8; s <= 16; s *= 2) {
        const int npairs = 16 / s;
        const int stride = 2 * s * SM_LDD + 2 * s;  // from one pair's A to the next
        // T_p = B_p * Ainv_p -> Tmp (pair p at column offset p * s, rows 0 .. s)
        level_product(S + s * SM_LDD, SM_LDD, stride, S, SM_LDD, stride, sTmp, SM_LDT, s, s, npairs, 1.0);
        __syncthreads();
        // B_p = -Cinv_p * T_p
        level_product(S + s * SM_LDD + s, SM_LDD, stride, sTmp, SM_LDT, s, S + s * SM_LDD, SM_LDD, stride, s, npairs, -1.0);
        __syncthreads();
    }
}

// In-place Cholesky of the lower triangle of the padded n x n matrix Mat (n = 32 nb), real dimension T, left-looking
// in 32-column block steps.  If Vt != nullptr the diagonal tiles of Vt = C^-T (upper) and the inverted diagonal
// blocks Dinv [nb][32][32] are written as well.  Ends with a barrier: every write is visible to the whole CTA.
__device__ __forceinline__ void potrf_small(double* Mat, int ld, int nb, int T, double* Vt, double* Dinv, int* info_b,
                                            const SmallSmem& sm) {
    const int warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int k = 0; k < nb; ++k) {
        const int kc = SB * k;
        const int nr = max(0, min(SB, T - kc));
        if (k > 0) load_panel(sm.B, Mat + (size_t)kc * ld, ld, kc);
        __syncthreads();
        // column update of every 16-row sub-block at or below the diagonal block
#pragma unroll 1
        for (int r = first_owned(2 * k, warp); r < 2 * nb; r += SM_WARPS) {
            SubAcc acc, c;
            sub_zero(acc);
            sub_gemm(acc, Mat + (size_t)16 * r * ld, ld, sm.B, 0, 0, kc);
            sub_load(c, Mat + (size_t)16 * r * ld + kc, ld);
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    c.v[mi][ni][0] -= acc.v[mi][ni][0];
                    c.v[mi][ni][1] -= acc.v[mi][ni][1];
                }
            if (r < 2 * k + 2)
                sub_store(c, sm.D + (r - 2 * k) * 16 * SM_LDD, SM_LDD);
            else
                sub_store(c, Mat + (size_t)16 * r * ld + kc, ld);  // unscaled; scaled below once the inverse exists
        }
        __syncthreads();
        {
            const int bad = diag_factor32(sm.D, sm.rd, nr, kc);
            if (threadIdx.x == 0 && bad != 0 && info_b && *info_b == 0) *info_b = bad;
            diag_inverse32(sm.D, sm.rd, sm.Inv, sm.Tmp, nr);
        }
        // the factor's diagonal block (lower, real columns only; strict upper part of the real rows cleaned), the
        // inverse for the triangular-inverse stage and the diagonal tile of V = C^-T
        for (int e = threadIdx.x; e < SB * SB; e += SM_THREADS) {
            const int r = e >> 5, c = e & 31;
            double* dst = Mat + (size_t)(kc + r) * ld + kc + c;
            if (c < nr && c <= r)
                *dst = sm.D[r * SM_LDD + c];
            else if (r < nr && c > r)
                *dst = 0.0;
            if (Dinv) Dinv[(size_t)k * SB * SB + e] = c <= r ? sm.Inv[r * SM_LDD + c] : 0.0;
            if (Vt) Vt[(size_t)(kc + r) * ld + kc + c] = r <= c ? sm.Inv[c * SM_LDD + r] : 0.0;
        }
        // panel below the diagonal block: L[r, k] = unscaled * Linv^T
#pragma unroll 1
        for (int r = first_owned(2 * k + 2, warp); r < 2 * nb; r += SM_WARPS) {
            SubAcc t, x;
            sub_load(t, Mat + (size_t)16 * r * ld + kc, ld);
            sub_scale(x, t, sm.Inv, 1.0);
            sub_store(x, Mat + (size_t)16 * r * ld + kc, ld);
        }
        __syncthreads();  // the next step's panel is made of rows written here
    }
}

// V = C^-T (upper, row-major) from the factor C (lower, in Cm) and the inverted diagonal blocks; the diagonal tiles
// of V are already in place.  Column step i:  V[j, i] = -(sum_{p=j}^{i-1} V[j, p] C[i, p]^T) Dinv[i]^T,  j < i.
__device__ __forceinline__ void trtri_small(double* Vt, const double* Cm, int ld, int nb, const double* Dinv,
                                            const SmallSmem& sm) {
    const int warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int i = 1; i < nb; ++i) {
        const int ic = SB * i;
        // the inverted diagonal block rides in the same cp.async group as the panel
        for (int c = threadIdx.x; c < SB * SB / 2; c += SM_THREADS)
            cp_async16(sm.Inv + (c >> 4) * SM_LDD + 2 * (c & 15), Dinv + (size_t)i * SB * SB + 2 * c, true);
        load_panel(sm.B, Cm + (size_t)ic * ld, ld, ic);
        __syncthreads();
        // row sub-block r contracts K = 32 (i - r / 2) columns: warps take r and (2 i - 1 - r) in turns ("snake"), so
        // that every warp's total K is about the same
#pragma unroll 1
        for (int t = warp; t < 2 * i; t += SM_WARPS) {
            const int r = t >= SM_WARPS ? 2 * i - 1 - (t - SM_WARPS) : t;  // at most two rounds (2 i <= 16)
            SubAcc acc, x;
            sub_zero(acc);
            sub_gemm(acc, Vt + (size_t)16 * r * ld, ld, sm.B, 0, SB * (r >> 1), ic);
            sub_scale(x, acc, sm.Inv, -1.0);
            sub_store(x, Vt + (size_t)16 * r * ld + ic, ld);
        }
        __syncthreads();
    }
}

// ---- kernel A: M = C C^T, u = C^-1 y (row T), V = C^-T ------------------------------------------------------
// Mat [B][n][n] holds the lower tiles of M = K + diag(y_var + jitter) with row T = y_mean (k_matern32<1>); on return
// it holds C (row T zeroed), Vt holds C^-T, u [B][T] the forward-substituted right-hand side.
__global__ void __launch_bounds__(SM_THREADS, 2)
    k_small_factor_inverse(double* Mat, double* Vt, double* Dinv, double* u, int* info, int n, int T) {
    extern __shared__ __align__(16) double small_smem[];
    const SmallSmem sm(small_smem);
    const int b = blockIdx.x, nb = n / SB;
    double* Mb = Mat + (size_t)b * n * n;
    double* Vb = Vt + (size_t)b * n * n;
    double* Db = Dinv + (size_t)b * nb * SB * SB;
    potrf_small(Mb, n, nb, T, Vb, Db, info + b, sm);
    for (int j = threadIdx.x; j < T; j += SM_THREADS) {
        double* p = Mb + (size_t)T * n + j;
        u[(size_t)b * T + j] = *p;
        *p = 0.0;  // the triangular inverse sees blockdiag(C, I)
    }
    __syncthreads();
    trtri_small(Vb, Mb, n, nb, Db, sm);
}

// ---- kernel B: cov = D + E - E (V V^T) E, then scale_tri = chol(cov) with (1, mu) riding along ---------------------
// Epilogue semantics are those of k_lauum_cov (be_kernels.cuh): Work (aliases the buffer that held C) receives the
// padded lower covariance, rows T / T+1 = (1, mu), identity padding; var_diag and the optional dense symmetric
// covariance are written on the way.
__global__ void __launch_bounds__(SM_THREADS, 2)
    k_small_cov_factor(const double* Vt, double* Work, const double* __restrict__ y_var, double jitter,
                       const double* __restrict__ mu, double* __restrict__ var_diag, double* __restrict__ cov_dense,
                       int* info, int n, int T) {
    extern __shared__ __align__(16) double small_smem[];
    const SmallSmem sm(small_smem);
    const int b = blockIdx.x, nb = n / SB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const double* Vb = Vt + (size_t)b * n * n;
    double* Wb = Work + (size_t)b * n * n;
    const double* yv = y_var + (size_t)b * T;
    const double* mub = mu + (size_t)b * T;
    double* cd = cov_dense ? cov_dense + (size_t)b * T * T : nullptr;
#pragma unroll 1
    for (int j = 0; j < nb; ++j) {
        const int jc = SB * j;
        load_panel(sm.B, Vb + (size_t)jc * n + jc, n, n - jc);
        __syncthreads();
#pragma unroll 1
        for (int r = first_owned(2 * j, warp); r < 2 * nb; r += SM_WARPS) {
            SubAcc acc;
            sub_zero(acc);
            sub_gemm(acc, Vb + (size_t)16 * r * n, n, sm.B, jc, SB * (r >> 1), n);
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const int gr = 16 * r + 8 * mi + g;
                const double dr = gr < T ? yv[gr] : 0.0;
                const double er = dr + jitter;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    double out[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int gc = jc + 8 * ni + 2 * q + e;
                        double val;
                        if (gr < T && gc < T) {
                            const double ec = yv[gc] + jitter;
                            val = -er * ec * acc.v[mi][ni][e];
                            if (gr == gc) {
                                val += dr + er;
                                var_diag[(size_t)b * T + gr] = val;
                            }
                            if (cd && gc <= gr) {
                                cd[(size_t)gr * T + gc] = val;
                                cd[(size_t)gc * T + gr] = val;
                            }
                        } else if (gr == T && gc < T) {
                            val = 1.0;
                        } else if (gr == T + 1 && gc < T) {
                            val = mub[gc];
                        } else {
                            val = gr == gc ? 1.0 : 0.0;
                        }
                        out[e] = val;
                    }
                    *reinterpret_cast<double2*>(Wb + (size_t)gr * n + jc + 8 * ni + 2 * q) = make_double2(out[0], out[1]);
                }
            }
        }
        __syncthreads();
    }
    potrf_small(Wb, n, nb, T, nullptr, nullptr, info + b, sm);
}

}  // namespace be
