// Kernels of the fit -> weight -> barycentre hot path (sm_100a, fp64).
//
// Internal matrix layout ("padded"): every T x T problem lives in a [Tp, ld] row-major buffer
// with Tp = ld = round_up(T + 2, 16).  Rows/cols >= T are padding: identity on the diagonal,
// zero elsewhere -- EXCEPT that rows T and T+1 (cols < T) may carry right-hand sides.  The
// blocked Cholesky never pivots on a padded column but applies every update to the padded
// rows, so after the factorisation those rows hold (L^-1 rhs)^T: forward substitution rides
// along inside the tensor-core updates instead of being a separate latency-bound TRSV.
#pragma once
#include <math.h>

#include "dmma_gemm.cuh"
#include "chol_diag.cuh"

namespace be {

constexpr int NB = 128;  // block size of the blocked algorithms == GEMM tile edge
constexpr int DIAG_SMEM_BYTES = DG_SMEM_BYTES;
constexpr double SQRT3 = 1.7320508075688772;
constexpr double LOG_2PI = 1.8378770664093453;

__host__ __device__ inline int pad_dim(int T) { return ((T + 2 + 15) / 16) * 16; }
__host__ __device__ inline int num_blocks(int Tp) { return (Tp + NB - 1) / NB; }
__device__ __forceinline__ int blk_rows(int Tp, int kb) { return min(NB, Tp - kb * NB); }

// --------------------------------------------------------------------------------------
// models.py:175-182: X = realisation_set.T, y_mean (arithmetic), y_var = np.var(axis=0)
// --------------------------------------------------------------------------------------
__global__ void k_gpdtw1d_inputs(const double* __restrict__ reals, int B, int R, int T, double* __restrict__ X,
                                 double* __restrict__ y_mean, double* __restrict__ y_var) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)B * T) return;
    int b = (int)(gid / T), t = (int)(gid % T);
    const double* src = reals + (size_t)b * R * T + t;
    double s = 0.0;
    for (int r = 0; r < R; ++r) s += src[(size_t)r * T];
    double mean = s / R;
    double v = 0.0;
    for (int r = 0; r < R; ++r) {
        double x = src[(size_t)r * T];
        double d = x - mean;
        v += d * d;
        if (X) X[((size_t)b * T + t) * R + r] = x;
    }
    if (y_mean) y_mean[gid] = mean;
    if (y_var) y_var[gid] = v / R;
}

// --------------------------------------------------------------------------------------
// Matern-3/2 gram (gpflow Matern32 on X/l with the |x|^2+|y|^2-2xy expansion, r2 clamped at
// 1e-36).  One CTA per 128 x 128 tile.  MODE 0: dense symmetric K [B,T,T].
// MODE 1: padded lower tiles of M = K + diag(y_var + jitter), padding identity, row T = y_mean.
// MODE 2: padded [Tp,ld] K, all tiles, no noise, zero padding (predict_f's Kmn, models.py:217).
// --------------------------------------------------------------------------------------
// sqrt(q) for a NORMAL positive q (callers clamp at 1e-36): MUFU seed, one Newton step on 1/sqrt and
// one on the root -- correctly rounded in all but a vanishing share of cases, no special-case calls.
__device__ __forceinline__ double sqrt_pos(double q) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
    double e = fma(-q * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    double r = q * y;
    return fma(fma(-r, r, q), 0.5 * y, r);
}

// exp(-x) for x >= 0: n = rint(-x log2 e) by the 1.5 * 2^52 trick, Cody-Waite reduction with a
// two-part ln 2, degree-13 Taylor polynomial on |r| <= ln2 / 2 (truncation 4e-18), exponent patched
// in as an integer.  Valid for 0 <= x <= 700 and branch-free, so that the compiler interleaves the
// chains of neighbouring entries; callers send x > 700 and NaN to the library routine.  <= 1 ulp from
// the library exp (tests/test_gpu_parity.py::test_matern32_gram holds 1e-12 vs NumPy).
__device__ __forceinline__ double exp_core(double y) {  // exp(y), |y| <= 700
    const double MAGIC = 6755399441055744.0;
    double t = fma(y, 1.4426950408889634, MAGIC);
    int n = __double2loint(t);
    double nf = t - MAGIC;
    double r = fma(nf, -6.93147180369123816490e-01, y);
    r = fma(nf, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;           // 1/13!
    p = fma(p, r, 2.08767569878681e-09);         // 1/12!
    p = fma(p, r, 2.505210838544172e-08);        // 1/11!
    p = fma(p, r, 2.755731922398589e-07);        // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);       // 1/9!
    p = fma(p, r, 2.48015873015873e-05);         // 1/8!
    p = fma(p, r, 1.984126984126984e-04);        // 1/7!
    p = fma(p, r, 1.388888888888889e-03);        // 1/6!
    p = fma(p, r, 8.333333333333333e-03);        // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);       // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);       // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return p * __hiloint2double((1023 + n) << 20, 0);
}
__device__ __forceinline__ double exp_neg(double x) { return exp_core(-x); }

// Thread t = (ty, tx) = (t / 16, t % 16) owns rows ty*8 .. ty*8+7 and the column pairs
// (2 tx, 2 tx + 1) + 32 c, c = 0..3, of the tile: the scaled inputs sit k-major in shared memory
// ([k][row]) so that every operand fetch is a conflict-free LDS.128 (broadcast for the rows) and
// every store is a 16-byte one with 16 lanes covering 256 contiguous bytes of a row.  What remains
// per entry is the FP64 sqrt and exp, which is what bounds this kernel (FP64 pipe, not HBM).
#ifndef BE_MATERN_CTAS
#define BE_MATERN_CTAS 4
#endif
template <int MODE>
__global__ void __launch_bounds__(256, BE_MATERN_CTAS) k_matern32(const double* __restrict__ X, int B, int T, int R,
                                                     const double* __restrict__ variance,
                                                     const double* __restrict__ lengthscale,
                                                     const double* __restrict__ y_mean, const double* __restrict__ y_var,
                                                     double jitter, double* __restrict__ out, int Tp, int ld,
                                                     int ntiles) {
    extern __shared__ __align__(16) double sm[];
    double* xi = sm;              // [R][128]  rows of the tile, k-major
    double* xj = xi + NB * R;     // [R][128]  columns of the tile
    double* si = xj + NB * R;     // |x_i / l|^2
    double* sj = si + NB;
    int tile = blockIdx.x / B, b = blockIdx.x % B;
    int ti, tj;
    if (MODE == 1) {
        tri_decode(tile, ti, tj);
    } else {
        int nt = ((MODE == 2 ? Tp : T) + NB - 1) / NB;
        ti = tile / nt;
        tj = tile % nt;
    }
    const double ls = lengthscale[b], var = variance[b];
    const double* Xb = X + (size_t)b * T * R;
    for (int e = threadIdx.x; e < NB * R; e += blockDim.x) {
        int r = e / R, k = e % R;
        int gi = ti * NB + r, gj = tj * NB + r;
        xi[k * NB + r] = gi < T ? Xb[(size_t)gi * R + k] / ls : 0.0;
        xj[k * NB + r] = gj < T ? Xb[(size_t)gj * R + k] / ls : 0.0;
    }
    __syncthreads();
    {
        const double* src = threadIdx.x < NB ? xi : xj;
        const int r = threadIdx.x & (NB - 1);
        double s = 0.0;
        for (int k = 0; k < R; ++k) s += src[k * NB + r] * src[k * NB + r];
        (threadIdx.x < NB ? si : sj)[r] = s;
    }
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int lim = MODE == 0 ? T : Tp;
    // fast path: every entry of the tile is a real (i, j) pair and (MODE 1) no diagonal entry is in it
    const bool interior = (ti + 1) * NB <= T && (tj + 1) * NB <= T && !(MODE == 1 && ti == tj);
    double* ob = MODE != 0 ? out + (size_t)b * Tp * ld : out + (size_t)b * T * T;
    const int ldo = MODE != 0 ? ld : T;
    // 2 x 2 entries at a time (two rows, one column pair): few registers, so four CTAs (32 warps) stay
    // resident per SM and hide the long dependent chains of the FP64 sqrt and exp.
#pragma unroll 1
    for (int ip = 0; ip < 4; ++ip) {
        const int lr = ty * 8 + 2 * ip;  // local rows lr, lr + 1
        const int gi0 = ti * NB + lr;
        const double2 si2 = *reinterpret_cast<const double2*>(si + lr);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            // MODE 1 writes the LOWER triangle: in a diagonal tile the 32-column groups strictly right of the rows'
            // own 32-block are never read (the factorisations load j <= i, the small-T kernels 32-blocks at or below
            // the diagonal) -- 6 of the tile's 16 sub-blocks, a quarter of the whole gram at T = 251
            if (MODE == 1 && ti == tj && c > (lr >> 5)) continue;
            const int lc = 2 * tx + 32 * c;  // local columns lc, lc + 1
            const int gj = tj * NB + lc;
            double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            for (int k = 0; k < R; ++k) {
                const double2 a2 = *reinterpret_cast<const double2*>(xi + k * NB + lr);
                const double2 b2 = *reinterpret_cast<const double2*>(xj + k * NB + lc);
                acc[0][0] = fma(a2.x, b2.x, acc[0][0]);
                acc[0][1] = fma(a2.x, b2.y, acc[0][1]);
                acc[1][0] = fma(a2.y, b2.x, acc[1][0]);
                acc[1][1] = fma(a2.y, b2.y, acc[1][1]);
            }
            const double2 sj2 = *reinterpret_cast<const double2*>(sj + lc);
            double v[2][2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double r2 = (-2.0 * acc[i][e] + (i ? si2.y : si2.x)) + (e ? sj2.y : sj2.x);
                    const double x = SQRT3 * sqrt_pos(fmax(r2, 1e-36));
                    v[i][e] = x;
                }
            const bool fast = v[0][0] <= 700.0 && v[0][1] <= 700.0 && v[1][0] <= 700.0 && v[1][1] <= 700.0;
            if (fast) {  // one straight-line block: the four polynomial chains interleave
                const double e00 = exp_neg(v[0][0]), e01 = exp_neg(v[0][1]), e10 = exp_neg(v[1][0]), e11 = exp_neg(v[1][1]);
                v[0][0] = var * (1.0 + v[0][0]) * e00;
                v[0][1] = var * (1.0 + v[0][1]) * e01;
                v[1][0] = var * (1.0 + v[1][0]) * e10;
                v[1][1] = var * (1.0 + v[1][1]) * e11;
            } else {
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int e = 0; e < 2; ++e) v[i][e] = var * (1.0 + v[i][e]) * exp(-v[i][e]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int gi = gi0 + i;
                if (interior) {
                    if (MODE != 0) {
                        *reinterpret_cast<double2*>(ob + (size_t)gi * ldo + gj) = make_double2(v[i][0], v[i][1]);
                    } else {
                        ob[(size_t)gi * ldo + gj] = v[i][0];
                        ob[(size_t)gi * ldo + gj + 1] = v[i][1];
                    }
                    continue;
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int gc = gj + e;
                    if (gi >= lim || gc >= lim) continue;
                    double val;
                    if (gi < T && gc < T) {
                        val = v[i][e];
                        if (MODE == 1 && gi == gc) val += y_var[(size_t)b * T + gi] + jitter;
                    } else if (MODE == 1 && gi == T && gc < T) {
                        val = y_mean[(size_t)b * T + gc];
                    } else {
                        val = (MODE == 1 && gi == gc) ? 1.0 : 0.0;
                    }
                    ob[(size_t)gi * ldo + gc] = val;
                }
            }
        }
    }
}

// dense lower triangle -> padded buffer; rows T / T+1 optionally carry (1, mu)
__global__ void k_pad_from_dense(const double* __restrict__ A, const double* __restrict__ mu, int B, int T, int Tp,
                                 int ld, double* __restrict__ W, int with_rhs) {
    size_t n = (size_t)B * Tp * ld;
    for (size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gid < n; gid += (size_t)gridDim.x * blockDim.x) {
        int b = (int)(gid / ((size_t)Tp * ld));
        size_t rem = gid % ((size_t)Tp * ld);
        int i = (int)(rem / ld), j = (int)(rem % ld);
        double v;
        if (i < T && j < T)
            v = j <= i ? A[(size_t)b * T * T + (size_t)i * T + j] : 0.0;
        else if (with_rhs && i == T && j < T)
            v = 1.0;
        else if (with_rhs && i == T + 1 && j < T)
            v = mu[(size_t)b * T + j];
        else
            v = i == j ? 1.0 : 0.0;
        W[gid] = v;
    }
}

// --------------------------------------------------------------------------------------
// Tensor-core tile kernels (all share gemm_nt_mainloop)
// --------------------------------------------------------------------------------------

// Left-looking column update of the blocked Cholesky: for block rows ti >= kb
//   Mat[ti, kb] -= Mat[ti, 0:kb] * Mat[kb, 0:kb]^T        (one long-K GEMM per tile: K = kb * 128)
// Each output tile is read and written ONCE per factorisation (the right-looking form re-reads
// and re-writes it at every step with a K = 128 update, which left the tensor pipe ~55% idle).
// -C is preloaded into the accumulators so the read overlaps the pipeline prologue.
// The diagonal tile (ti == kb) is updated in place; the tiles below it go to the panel scratch
// Pbuf ([B][Tp][128], tile ti at rows ti*128) from which k_panel_scale writes the scaled panel
// back into Mat -- the scale cannot run in place because both 128 x 64 half-tile CTAs of a tile
// read the whole tile.
__global__ void __launch_bounds__(GEMM_THREADS, GEMM_CTAS_PER_SM)
    k_chol_update(double* __restrict__ Mat, double* __restrict__ Pbuf, int ld, int Tp, int kb, int B) {
    extern __shared__ __align__(16) double2 smem2[];
    int tile, half, b;  // 128 x 64 half-tiles: columns half*64 .. half*64+63 of block kb
    cta_decode(B, tile, half, b);
    const int ti = kb + tile;
    double* Mb = Mat + (size_t)b * Tp * ld;
    const int a_rows = blk_rows(Tp, ti), b_rows = min(BN, blk_rows(Tp, kb) - half * BN);
    if (b_rows <= 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* Cb = Mb + (size_t)ti * NB * ld + kb * NB + half * BN;
    TileAcc acc;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
        int r = acc_row(warp, lane, mi);
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            int c = acc_col(warp, lane, ni);
            double2 v = make_double2(0.0, 0.0);
            if (r < a_rows && c < b_rows) v = *reinterpret_cast<const double2*>(Cb + (size_t)r * ld + c);
            acc.v[mi][ni][0] = -v.x;
            acc.v[mi][ni][1] = -v.y;
        }
    }
    gemm_nt_mainloop<false>(Mb + (size_t)ti * NB * ld, ld, a_rows, Mb + (size_t)(kb * NB + half * BN) * ld, ld, b_rows,
                            kb * NB, smem2, acc);
    double* Ob = Cb;
    int ldo = ld;
    if (ti != kb) {
        Ob = Pbuf + ((size_t)b * Tp + (size_t)ti * NB) * NB + half * BN;
        ldo = NB;
    }
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
        int r = acc_row(warp, lane, mi);
        if (r >= a_rows) continue;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            int c = acc_col(warp, lane, ni);
            if (c >= b_rows) continue;
            *reinterpret_cast<double2*>(Ob + (size_t)r * ldo + c) = make_double2(-acc.v[mi][ni][0], -acc.v[mi][ni][1]);
        }
    }
}

// Right-multiplication of panel tiles (row block row_blk0 + x, column block cb) by the
// transposed inverse diagonal block:  Mat[ti, cb] <- sign * Pbuf[ti] * Dinv[cb]^T
// (panel TRSM of the Cholesky with sign=+1; second half of the trtri column step with -1).
// The unscaled tiles live in the panel scratch Pbuf ([B][Tp][128]).
__global__ void __launch_bounds__(GEMM_THREADS, GEMM_CTAS_PER_SM)
    k_panel_scale(const double* __restrict__ Pbuf, double* __restrict__ Mat, int ld, int Tp, int row_blk0, int cb,
                  const double* __restrict__ Dinv, int nblk, double sign, int B) {
    extern __shared__ __align__(16) double2 smem2[];
    int tile, half, b;
    cta_decode(B, tile, half, b);
    int ti = row_blk0 + tile;
    double* Mb = Mat + (size_t)b * Tp * ld;
    const double* Sb = Pbuf + ((size_t)b * Tp + (size_t)ti * NB) * NB;
    const int a_rows = blk_rows(Tp, ti), kw = blk_rows(Tp, cb);
    const int b_rows = min(BN, kw - half * BN);
    if (b_rows <= 0) return;
    const double* Db = Dinv + ((size_t)b * nblk + cb) * NB * NB + (size_t)half * BN * NB;
    TileAcc acc;
    double* Cb = Mb + (size_t)ti * NB * ld + cb * NB + half * BN;
    // Dinv is lower triangular: output columns [half*64, half*64+64) only contract k < (half+1)*64
    const int klen = min(kw, (half + 1) * BN);
    gemm_nt_mainloop(Sb, NB, a_rows, Db, NB, b_rows, klen, smem2, acc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
        int r = acc_row(warp, lane, mi);
        if (r >= a_rows) continue;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            int c = acc_col(warp, lane, ni);
            if (c >= b_rows) continue;
            double2 v;
            v.x = sign * acc.v[mi][ni][0];
            v.y = sign * acc.v[mi][ni][1];
            *reinterpret_cast<double2*>(Cb + (size_t)r * ld + c) = v;
        }
    }
}

// trtri column step i, first half:  Pbuf[j] = sum_{p=j}^{i-1} V[j, p] * C[i, p]^T   (j < i)
// with V = C^-T (upper, row-major) so that every contraction is K-contiguous; k_panel_scale
// then writes V[j, i] = -Pbuf[j] * Dinv[i]^T.
__global__ void __launch_bounds__(GEMM_THREADS, GEMM_CTAS_PER_SM)
    k_trtri_accum(const double* __restrict__ V, const double* __restrict__ Cm, double* __restrict__ Pbuf, int ld,
                  int Tp, int i, int B) {
    extern __shared__ __align__(16) double2 smem2[];
    int j, half, b;
    cta_decode(B, j, half, b);
    const double* Vb = V + (size_t)b * Tp * ld;
    const double* Cb = Cm + (size_t)b * Tp * ld;
    const int b_rows = min(BN, blk_rows(Tp, i) - half * BN);
    if (b_rows <= 0) return;
    TileAcc acc;
    gemm_nt_mainloop(Vb + (size_t)j * NB * ld + j * NB, ld, NB, Cb + (size_t)(i * NB + half * BN) * ld + j * NB, ld,
                     b_rows, (i - j) * NB, smem2, acc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* Ob = Pbuf + ((size_t)b * Tp + (size_t)j * NB) * NB + half * BN;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
        int r = acc_row(warp, lane, mi);
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            int c = acc_col(warp, lane, ni);
            if (c >= b_rows) continue;
            *reinterpret_cast<double2*>(Ob + (size_t)r * NB + c) = make_double2(acc.v[mi][ni][0], acc.v[mi][ni][1]);
        }
    }
}

// lauum with the posterior epilogue.  Minv = V V^T (V = C^-T, C = chol(K + E), E = D + jitter I);
//   cov = D + E - E Minv E            ( == K - K (K+E)^-1 K + D, see DESIGN.md )
// written to: Work (padded, lower; aliases the buffer that held C), var_diag, and optionally
// the dense symmetric cov [B,T,T].  Rows T / T+1 of Work get the right-hand sides (1, mu) of
// the log-likelihood stage; the rest of the padding is identity.
__global__ void __launch_bounds__(GEMM_THREADS, GEMM_CTAS_PER_SM)
    k_lauum_cov(const double* __restrict__ V, int ld, int Tp, int T, const double* __restrict__ y_var, double jitter,
                const double* __restrict__ mu, double* __restrict__ Work, double* __restrict__ var_diag,
                double* __restrict__ cov_dense, int B, int row_major_raster) {
    extern __shared__ __align__(16) double2 smem2[];
    int tile, half, b;
    int ti, tj;
    if (row_major_raster) {
        // Block-row-major order: all (problem, tj, half) of block row ti side by side.  Every CTA of the group has the
        // same K length (Tp - 128 ti, as in the tile-major order, so waves still drain together) and CONSECUTIVE CTAs
        // are the 2 (ti + 1) half-tiles of ONE problem's block row: they read the same A operand (V rows ti) through
        // L2 instead of each from DRAM (tile-major: 12.8x the algorithmic bytes, profiles/traffic.json r01).
        const long long idx = blockIdx.x;
        int t = (int)((sqrt(1.0 + 4.0 * (double)idx / (double)B) - 1.0) * 0.5);
        while ((long long)B * (t + 1) * (t + 2) <= idx) ++t;
        while ((long long)B * t * (t + 1) > idx) --t;
        ti = t;
        const int rem = (int)(idx - (long long)B * t * (t + 1));
        b = rem / (2 * (ti + 1));
        const int r2 = rem - b * 2 * (ti + 1);
        tj = r2 >> 1;
        half = r2 & 1;
    } else {
        cta_decode(B, tile, half, b);
        tri_decode(tile, ti, tj);
    }
    const double* Vb = V + (size_t)b * Tp * ld;
    const int a_rows = blk_rows(Tp, ti), b_rows = min(BN, blk_rows(Tp, tj) - half * BN);
    if (b_rows <= 0) return;
    TileAcc acc;
    gemm_nt_mainloop(Vb + (size_t)ti * NB * ld + ti * NB, ld, a_rows, Vb + (size_t)(tj * NB + half * BN) * ld + ti * NB,
                     ld, b_rows, Tp - ti * NB, smem2, acc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* Wb = Work + (size_t)b * Tp * ld;
    const double* yv = y_var + (size_t)b * T;
    const double* mub = mu + (size_t)b * T;
    double* cd = cov_dense ? cov_dense + (size_t)b * T * T : nullptr;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
        int r = acc_row(warp, lane, mi);
        if (r >= a_rows) continue;
        int gr = ti * NB + r;
        double dr = gr < T ? yv[gr] : 0.0;
        double er = dr + jitter;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            int c = acc_col(warp, lane, ni);
            if (c >= b_rows) continue;
            double out[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int gc = tj * NB + half * BN + c + e;
                double val;
                if (gr < T && gc < T) {
                    double ec = yv[gc] + jitter;
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
            *reinterpret_cast<double2*>(Wb + (size_t)gr * ld + tj * NB + half * BN + c) = make_double2(out[0], out[1]);
        }
    }
}

// --------------------------------------------------------------------------------------
// small memory-bound helpers
// --------------------------------------------------------------------------------------

// copy padded row `row` (cols < T) to out[b, 0:T] and zero it in the matrix
__global__ void k_extract_row(double* __restrict__ Mat, int ld, int Tp, int T, int row, double* __restrict__ out,
                              int B, int zero_after) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)B * T) return;
    int b = (int)(gid / T), j = (int)(gid % T);
    double* p = Mat + (size_t)b * Tp * ld + (size_t)row * ld + j;
    out[gid] = *p;
    if (zero_after) *p = 0.0;
}

// mean = y - E * (V u),  u = C^-1 y (rode along as row T of the first factorisation).
// One warp per row: alpha_i = sum_{k>=i} V[i,k] u[k]   (V upper, K-contiguous => coalesced).
__global__ void __launch_bounds__(256) k_posterior_mean(const double* __restrict__ V, int ld, int Tp, int T,
                                                        const double* __restrict__ u, const double* __restrict__ y_mean,
                                                        const double* __restrict__ y_var, double jitter,
                                                        double* __restrict__ mu, int B,
                                                        double* __restrict__ var_diag = nullptr) {
    int warp_global = blockIdx.x * 8 + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (warp_global >= B * T) return;
    int b = warp_global / T, i = warp_global % T;
    const double* row = V + (size_t)b * Tp * ld + (size_t)i * ld;
    const double* ub = u + (size_t)b * T;
    double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
    int k0 = i & ~1;  // aligned start; element k0 < i (if any) is an explicit zero of the upper factor
    for (int k = k0 + 2 * lane; k < T; k += 64) {
        double2 v = *reinterpret_cast<const double2*>(row + k);
        if (k >= i) {
            s0 += v.x * ub[k];
            q0 = fma(v.x, v.x, q0);
        }
        if (k + 1 < T && k + 1 >= i) {
            s1 += v.y * ub[k + 1];
            q1 = fma(v.y, v.y, q1);
        }
    }
    double s = s0 + s1, q = q0 + q1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
        size_t g = (size_t)b * T + i;
        const double E = y_var[g] + jitter;
        mu[g] = y_mean[g] - E * s;
        // factored mode: cov_ii = D + E - E^2 (M^-1)_ii with (M^-1)_ii = |row i of V|^2
        if (var_diag) var_diag[g] = (y_var[g] + E) - E * E * q;
    }
}

// ---- factored posterior (be_gp_posterior_factored): cov = E' - E M^-1 E is never formed.  Woodbury:
//   cov^-1 = E'^-1 + G N^-1 G,  det cov = det N det E' / det M,  E' = D + E,  G = E / E',  N = K + diag(E D / E').
// k_factored_prepare: per problem the diagonal n = E D / E' of N, the right-hand sides g = G 1 and g*mu that
// ride along in chol(N), and the diagonal parts of the four statistics:
//   base = (sum 1/E', sum mu/E', sum mu^2/E', 1/2 sum log E').  One CTA per problem.
__global__ void __launch_bounds__(256) k_factored_prepare(const double* __restrict__ y_var, const double* __restrict__ mu,
                                                          double jitter, int T, double* __restrict__ nvar,
                                                          double* __restrict__ g, double* __restrict__ gmu,
                                                          double* __restrict__ base) {
    __shared__ double red[4][8];
    const int b = blockIdx.x;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j = threadIdx.x; j < T; j += 256) {
        const size_t o = (size_t)b * T + j;
        const double D = y_var[o], E = D + jitter, Ep = D + E, m = mu[o];
        nvar[o] = E * D / Ep;
        const double gg = E / Ep;
        g[o] = gg;
        gmu[o] = gg * m;
        v[0] += 1.0 / Ep;
        v[1] += m / Ep;
        v[2] += m * m / Ep;
        v[3] += 0.5 * log(Ep);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        base[(size_t)b * 4 + threadIdx.x] = s;
    }
}

// padded row `row` (cols < T) := src[b, 0:T]
__global__ void k_set_row(double* __restrict__ Mat, int ld, int Tp, int T, int row, const double* __restrict__ src, int B) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)B * T) return;
    int b = (int)(gid / T), j = (int)(gid % T);
    Mat[(size_t)b * Tp * ld + (size_t)row * ld + j] = src[gid];
}

// mvn_stats = base + (|a'|^2, a'.b', |b'|^2, sum log diag L_N) - (0, 0, 0, sum log diag C)
__global__ void k_factored_finish(const double* __restrict__ statsN, const double* __restrict__ base,
                                  const double* __restrict__ statsC, int B, double* __restrict__ stats) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= B * 4) return;
    double v = base[gid] + statsN[gid];
    if ((gid & 3) == 3) v -= statsC[gid];
    stats[gid] = v;
}

// mvn_stats[b] = (|a|^2, a.b, |b|^2, sum log diag L) from rows T, T+1 and the diagonal
__global__ void __launch_bounds__(256) k_mvn_stats(const double* __restrict__ Work, int ld, int Tp, int T,
                                                   double* __restrict__ stats) {
    __shared__ double red[4][8];
    int b = blockIdx.x;
    const double* Wb = Work + (size_t)b * Tp * ld;
    const double* ra = Wb + (size_t)T * ld;
    const double* rb = Wb + (size_t)(T + 1) * ld;
    double aa = 0, ab = 0, bb = 0, ld_ = 0;
    for (int j = threadIdx.x; j < T; j += 256) {
        double a = ra[j], bq = rb[j];
        aa += a * a;
        ab += a * bq;
        bb += bq * bq;
        ld_ += log(fabs(Wb[(size_t)j * ld + j]));
    }
    double v[4] = {aa, ab, bb, ld_};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        stats[(size_t)b * 4 + threadIdx.x] = s;
    }
}

// padded lower factor -> dense [B,T,T] with the strict upper triangle zeroed
__global__ void k_copy_out_tri(const double* __restrict__ Work, int ld, int Tp, int T, double* __restrict__ L, int B) {
    size_t n = (size_t)B * T * T;
    for (size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gid < n; gid += (size_t)gridDim.x * blockDim.x) {
        int b = (int)(gid / ((size_t)T * T));
        size_t rem = gid % ((size_t)T * T);
        int i = (int)(rem / T), j = (int)(rem % T);
        L[gid] = j <= i ? Work[(size_t)b * Tp * ld + (size_t)i * ld + j] : 0.0;
    }
}

__global__ void k_diag_from_dense(const double* __restrict__ A, int B, int T, double* __restrict__ d) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)B * T) return;
    int b = (int)(gid / T), i = (int)(gid % T);
    d[gid] = A[(size_t)b * T * T + (size_t)i * T + i];
}

// --------------------------------------------------------------------------------------
// a4: log-likelihood weights.  weights.py:97-100 evaluates the MVN density at the CONSTANT
// vector o*1_T (quirk Q-LL):  -1/2 |L^-1 (o 1 - mu)|^2 = -1/2 (o^2 |a|^2 - 2 o a.b + |b|^2).
// --------------------------------------------------------------------------------------
__device__ __forceinline__ double constvec_ll(const double* __restrict__ st, double o, int T) {
    double maha = (o * o) * st[0] - 2.0 * o * st[1] + st[2];
    return -0.5 * maha - 0.5 * (double)T * LOG_2PI - st[3];
}

__global__ void k_mvn_constvec_logprob(const double* __restrict__ stats, const double* __restrict__ obs, int C, int M,
                                       int Ro, int T, double* __restrict__ ll) {
    size_t n = (size_t)C * M * Ro * T;
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n) return;
    int i = (int)(gid % T);
    size_t q = gid / T;
    int r = (int)(q % Ro);
    q /= Ro;
    int m = (int)(q % M);
    int c = (int)(q / M);
    ll[gid] = constvec_ll(stats + ((size_t)c * M + m) * 4, obs[((size_t)c * Ro + r) * T + i], T);
}

// distrax MultivariateNormalTri.log_prob for GENERAL vectors (dists.py; the NLL of utils.py:139): one CTA per
// vector x, z = L^-1 (x - mu) by blocked forward substitution against the STORED dense factor scale_tri [T,T]
// (32 columns at a time: warp 0 solves the 32 x 32 diagonal block with shuffles, then every thread takes the rows
// below it), ll = -1/2 |z|^2 - T/2 log 2pi - sum log diag L.  O(T^2) per vector, L is read once per vector.
__global__ void __launch_bounds__(256) k_mvn_logprob_vectors(const double* __restrict__ mu, const double* __restrict__ L,
                                                              const double* __restrict__ x, int T, int N,
                                                              double sum_log_diag, double* __restrict__ ll) {
    extern __shared__ double zsh[];  // [T] running right-hand side, then the solution
    __shared__ double red[8];
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < T; i += 256) zsh[i] = x[(size_t)v * T + i] - mu[i];
    __syncthreads();
    for (int j0 = 0; j0 < T; j0 += 32) {
        const int w = min(32, T - j0);
        if (warp == 0) {
            // lane l owns row j0 + l of the diagonal block
            double r = lane < w ? zsh[j0 + lane] : 0.0;
            const double* row = L + (size_t)(j0 + min(lane, w - 1)) * T + j0;
            for (int k = 0; k < w; ++k) {
                if (lane == k) r = r / row[k];  // z_k is final
                const double zk = __shfl_sync(0xffffffffu, r, k);
                if (lane > k && lane < w) r = fma(-row[k], zk, r);
            }
            if (lane < w) zsh[j0 + lane] = r;
        }
        __syncthreads();
        for (int i = j0 + 32 + tid; i < T; i += 256) {
            const double* row = L + (size_t)i * T + j0;
            double acc = 0.0;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) acc = fma(row[k], zsh[j0 + k], acc);
            zsh[i] -= acc;
        }
        __syncthreads();
    }
    double q = 0.0;
    for (int i = tid; i < T; i += 256) q = fma(zsh[i], zsh[i], q);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) red[warp] = q;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += red[k];
        ll[v] = -0.5 * s - 0.5 * (double)T * LOG_2PI - sum_log_diag;
    }
}

// Un-normalised weights of one point are staged in SHARED memory (one column per thread, [M][blockDim.x],
// conflict-free) between the pass that forms them and the pass that divides by their sum, so that the
// weights cross HBM once (the first version wrote them, re-read them and wrote them again: 2.4x the
// algorithmic bytes at M = 24).  When M * blockDim.x * 8 bytes do not fit the launch passes smem_ok = 0 and
// the output array itself is the staging buffer, as before.
struct WeightStage {
    double* p;
    size_t stride;
    __device__ __forceinline__ WeightStage(double* smem, bool smem_ok, double* w_point, size_t w_stride) {
        p = smem_ok ? smem + threadIdx.x : w_point;
        stride = smem_ok ? blockDim.x : w_stride;
    }
    __device__ __forceinline__ double& operator[](int m) { return p[(size_t)m * stride]; }
};
constexpr size_t WEIGHT_STAGE_MAX_BYTES = 96 * 1024;
__host__ inline int weight_stage_block(int M) { return M <= 32 ? 128 : 64; }
__host__ inline size_t weight_stage_bytes(int M) {
    size_t b = (size_t)M * weight_stage_block(M) * sizeof(double);
    return b <= WEIGHT_STAGE_MAX_BYTES ? b : 0;
}
// k_loglik_weights_mvn also keeps the statistics of up to two cells ([2][M][4]) behind the staging area
__host__ inline size_t weight_stats_bytes(int M) { return (size_t)2 * M * 4 * sizeof(double); }


// one thread per (cell, time): mean over obs realisations (weights.py:103-104), exp(c .) (:107),
// normalise over models (:122-123).  The constant-vector log-density is quadratic in the observation,
// so its mean over the Ro realisations needs only mean(o) and mean(o^2):
//   mean_r ll = -1/2 (|a|^2 mean(o^2) - 2 a.b mean(o) + |b|^2) - T/2 log 2pi - sum log diag L
// (same terms as summing the Ro log-densities, one rounding pattern apart: ~1e-13 on the weights).
// That takes the kernel from M*Ro to M density evaluations per point (15% -> 42% of the HBM peak at 4M
// points, bench.py hbm_stages); the per-realisation values stay available from k_mvn_constvec_logprob.
// The (|a|^2, a.b, |b|^2, sum log diag L) statistics of the one or two cells a CTA's points belong to are
// copied to shared memory first (read per member from global memory they are a dependent L2 round trip per
// loop trip: the staging buffer leaves almost no L1).
// This form (library exp, M divisions) is ISSUE-bound: ~100 instructions per (point, member), 48 % of the HBM
// peak at 4 M points x 24 members.  It is what the library launches only when the staging area does not fit
// shared memory; otherwise k_loglik_weights_mvn_tab below (69 %).
__global__ void k_loglik_weights_mvn(const double* __restrict__ stats, const double* __restrict__ obs, int C, int M,
                                     int Ro, int T, double cst, double* __restrict__ w, double* __restrict__ lls_exp,
                                     double* __restrict__ lls_mean, int smem_ok) {
    extern __shared__ double wstage[];
    const size_t gid0 = (size_t)blockIdx.x * blockDim.x;
    const size_t gid = gid0 + threadIdx.x;
    const size_t n_pts = (size_t)C * T;
    const int c0 = (int)(gid0 / T);
    const int c_last = (int)((min(gid0 + blockDim.x, n_pts) - 1) / T);
    // smem_ok: bit 0 = the staging area exists, bit 1 = room for the statistics behind it
    double* sst = wstage + ((smem_ok & 1) ? (size_t)M * blockDim.x : 0);
    const bool stats_in_smem = (smem_ok & 2) && c_last - c0 <= 1;
    if (stats_in_smem) {
        const int n = (c_last - c0 + 1) * M * 4;
        for (int e = threadIdx.x; e < n; e += blockDim.x) sst[e] = stats[(size_t)c0 * M * 4 + e];
    }
    __syncthreads();
    if (gid >= n_pts) return;
    const int c = (int)(gid / T), i = (int)(gid - (size_t)c * T);
    WeightStage st(wstage, smem_ok & 1, w + (size_t)c * M * T + i, (size_t)T);
    const double* ob = obs + (size_t)c * Ro * T + i;
    double m1 = 0.0, m2 = 0.0;
#pragma unroll 5
    for (int r = 0; r < Ro; ++r) {
        double o = ob[(size_t)r * T];
        m1 += o;
        m2 = fma(o, o, m2);
    }
    m1 /= Ro;
    m2 /= Ro;
    const double base = -0.5 * (double)T * LOG_2PI;
    double total = 0.0;
    const double2* sp = stats_in_smem ? reinterpret_cast<const double2*>(sst + (size_t)(c - c0) * M * 4)
                                      : reinterpret_cast<const double2*>(stats + (size_t)c * M * 4);
    double* wp = w + (size_t)c * M * T + i;
#pragma unroll 4
    for (int m = 0; m < M; ++m) {
        const double2 s01 = sp[2 * m], s23 = sp[2 * m + 1];
        double maha = (m2 * s01.x - 2.0 * m1 * s01.y) + s23.x;
        double mean = (-0.5 * maha + base) - s23.y;
        double e = exp(cst * mean);
        size_t o = ((size_t)c * M + m) * T + i;
        if (lls_mean) lls_mean[o] = mean;
        if (lls_exp) lls_exp[o] = e;
        st[m] = e;
        if (e == e) total += e;  // xarray .sum('model') skips NaN (weights.py:122)
    }
#pragma unroll 4
    for (int m = 0; m < M; ++m) wp[(size_t)m * T] = st[m] / total;
}

// exp by a 16-entry table of 2^(j/16) and a degree-7 polynomial on |r| <= ln2/32 (truncation 1.2e-18, total
// error ~1.5 ulp): 13 FP64 instructions instead of the library's ~34 FP64 out of 67.  The table is 128 bytes
// of shared memory = each entry on its own pair of banks, so any index pattern is conflict-free.  Arguments
// outside |x| < 700 (overflow, gradual underflow, NaN) take the library's exp.
__constant__ double EXP2_16TH[16] = {
    0x1.0000000000000p+0, 0x1.0b5586cf9890fp+0, 0x1.172b83c7d517bp+0, 0x1.2387a6e756238p+0,
    0x1.306fe0a31b715p+0, 0x1.3dea64c123422p+0, 0x1.4bfdad5362a27p+0, 0x1.5ab07dd485429p+0,
    0x1.6a09e667f3bcdp+0, 0x1.7a11473eb0187p+0, 0x1.8ace5422aa0dbp+0, 0x1.9c49182a3f090p+0,
    0x1.ae89f995ad3adp+0, 0x1.c199bdd85529cp+0, 0x1.d5818dcfba487p+0, 0x1.ea4afa2a490dap+0};

// Constants whose low word is not zero are read as constant-bank operands of the DFMAs (as 64-bit literals
// they cost two moves each per use); ln2/16 is split so that its high part has a zero low word (an immediate).
struct ExpTabConsts {
    double inv, lo, c7, c6, c5, c4, c3;
};
__constant__ ExpTabConsts EXPC = {0x1.71547652b82fep+4,   0x1.fdf473de6af28p-26, 0x1.a01a01a01a01ap-13, 0x1.6c16c16c16c17p-10,
                                  0x1.1111111111111p-7,   0x1.5555555555555p-5,  0x1.5555555555555p-3};

// exp_tab16_core is branch-free (any input, garbage outside |x| < 700, never a trap or an out-of-range table
// index) so that the chains of neighbouring members interleave; exp_tab16_fix repairs the rare outliers.
__device__ __forceinline__ double exp_tab16_core(double x, const double* __restrict__ tab) {
    const double v = fma(x, EXPC.inv, 6755399441055744.0);  // k = rint(16 x / ln2) in the low word
    const int k = __double2loint(v);
    const double kd = v - 6755399441055744.0;
    double r = fma(kd, -0x1.62e4200000000p-5, x);
    r = fma(kd, -EXPC.lo, r);
    double q = fma(EXPC.c7, r, EXPC.c6);
    q = fma(q, r, EXPC.c5);
    q = fma(q, r, EXPC.c4);
    q = fma(q, r, EXPC.c3);
    q = fma(q, r, 0.5);
    const double p = fma(q * r, r, r);
    const double t = tab[k & 15];
    const double e = fma(t, p, t);
    return __hiloint2double(__double2hiint(e) + ((k >> 4) << 20), __double2loint(e));
}
__device__ __forceinline__ bool exp_tab16_ok(double x) { return fabs(x) < 700.0; }
__device__ __noinline__ double exp_tab16_fix(double x, double e) { return exp_tab16_ok(x) ? e : exp(x); }

// The form of k_loglik_weights_mvn the library launches when the staging area and the statistics both fit
// shared memory (smem layout: [M][blockDim] staging | [2][M][4] statistics | 16 table entries).  Same
// arithmetic up to the exponential; the exponential is exp_tab16 and the M divisions by the normaliser are
// one reciprocal and M products when the normaliser is an ordinary number (2^-1000 < total < 2^1000: the
// product is then within 1.5 ulp of the quotient); zero, subnormal, huge, infinite and NaN normalisers take
// the divisions, so 0/0, x/inf and NaN come out as in the reference (weights.py:122-123).
// 100 -> ~40 instructions per (point, member): the kernel was issue-bound (ncu r01n: 88 instructions per
// member executed, issue slots 68 % busy, DRAM at 35 %).
template <bool LLS>
__global__ void __launch_bounds__(128, 8)
    k_loglik_weights_mvn_tab(const double* __restrict__ stats, const double* __restrict__ obs, int C, int M, int Ro,
                             int T, double cst, double inv_ro, double* __restrict__ w, double* __restrict__ lls_exp,
                             double* __restrict__ lls_mean) {
    extern __shared__ double wstage[];
    // One 64-bit division per CTA-uniform quantity (gid0 / T); everything per thread is 32-bit and, when a row
    // is at least as long as the CTA (the usual case), division-free: the prologue was ~300 of the ~1400
    // instructions a warp executes at M = 24.
    const size_t gid0 = (size_t)blockIdx.x * blockDim.x;
    const size_t n_pts = (size_t)C * T;
    const unsigned uT = (unsigned)T;
    const int c0 = (int)(gid0 / uT);
    const unsigned i0 = (unsigned)(gid0 - (size_t)c0 * uT);
    const unsigned nvalid = (unsigned)min((size_t)blockDim.x, n_pts - gid0);
    const unsigned loc = i0 + threadIdx.x;
    unsigned spans, dc;
    if (uT >= blockDim.x) {  // the CTA's points lie in at most two cells
        spans = i0 + nvalid - 1 >= uT;
        dc = loc >= uT;
    } else {
        spans = (i0 + nvalid - 1) / uT;
        dc = loc / uT;
    }
    double* sst = wstage + (size_t)M * blockDim.x;
    double* tab = sst + (size_t)2 * M * 4;
    const bool stats_in_smem = spans <= 1;
    if (stats_in_smem) {
        const int n = (int)(spans + 1) * M * 4;
        const double* src = stats + (size_t)c0 * M * 4;
        for (int e = threadIdx.x; e < n; e += blockDim.x) sst[e] = src[e];
    }
    if (threadIdx.x < 16) tab[threadIdx.x] = EXP2_16TH[threadIdx.x];
    __syncthreads();
    if (threadIdx.x >= nvalid) return;
    const int c = c0 + (int)dc, i = (int)(loc - dc * uT);
    double* stg = wstage + threadIdx.x;
    const int bs = blockDim.x;
    const double* ob = obs + (size_t)c * Ro * T + i;
    double m1 = 0.0, m2 = 0.0;
#pragma unroll 5
    for (int r = 0; r < Ro; ++r) {
        double o = ob[(size_t)r * T];
        m1 += o;
        m2 = fma(o, o, m2);
    }
    m1 *= inv_ro;  // 1 / Ro from the host: within an ulp of the quotient the library-exp form takes
    m2 *= inv_ro;
    const double base = -0.5 * (double)T * LOG_2PI;
    const double m1x2 = 2.0 * m1;
    double total = 0.0;
    const double2* sp = stats_in_smem ? reinterpret_cast<const double2*>(sst + (size_t)dc * M * 4)
                                      : reinterpret_cast<const double2*>(stats + (size_t)c * M * 4);
    double* wp = w + (size_t)c * M * T + i;
    // members four at a time: the four exponential chains are independent and branch-free
    auto member = [&](int m, double& x, double& mean) {
        const double2 s01 = sp[2 * m], s23 = sp[2 * m + 1];
        double maha = (m2 * s01.x - m1x2 * s01.y) + s23.x;
        mean = (-0.5 * maha + base) - s23.y;
        x = cst * mean;
    };
    auto emit = [&](int m, double mean, double e) {
        if (LLS) {
            size_t o = ((size_t)c * M + m) * T + i;
            if (lls_mean) lls_mean[o] = mean;
            if (lls_exp) lls_exp[o] = e;
        }
        stg[m * bs] = e;
        if (e == e) total += e;  // xarray .sum('model') skips NaN (weights.py:122)
    };
    int m = 0;
    for (; m + 4 <= M; m += 4) {
        double x[4], mean[4], e[4];
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            member(m + j, x[j], mean[j]);
            e[j] = exp_tab16_core(x[j], tab);
            ok = ok && exp_tab16_ok(x[j]);
        }
        if (!ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = exp_tab16_fix(x[j], e[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) emit(m + j, mean[j], e[j]);
    }
    for (; m < M; ++m) {
        double x, mean;
        member(m, x, mean);
        emit(m, mean, exp_tab16_fix(x, exp_tab16_core(x, tab)));
    }
    if (total > 0x1p-1000 && total < 0x1p1000) {
        const double inv = 1.0 / total;
#pragma unroll 4
        for (int m = 0; m < M; ++m) wp[(size_t)m * T] = stg[m * bs] * inv;
    } else {
        for (int m = 0; m < M; ++m) wp[(size_t)m * T] = stg[m * bs] / total;
    }
}

// log by a 64-entry table: x = 2^e m, m in [1, 2); j = the top six mantissa bits, c_j = 1 + (j + 1/2) / 64 the centre of
// m's interval; r = m * (1 / c_j) - 1 (one FMA, |r| <= 2^-7); log x = e ln2 - log(1 / c_j) + log1p(r) with a degree-7
// polynomial (truncation < 2^-59).  The table holds the ROUNDED reciprocals and minus the logarithms of exactly those
// doubles, so the split is exact up to the final roundings: absolute error ~1.5e-16 (1 + |log x|).  ~20 instructions
// against the library's ~50; valid for positive NORMAL x (callers send everything else to the library).
__constant__ double LOG_INV_C[64] = {
    0x1.fc07f01fc07f0p-1, 0x1.f44659e4a4271p-1, 0x1.ecc07b301ecc0p-1, 0x1.e573ac901e574p-1,
    0x1.de5d6e3f8868ap-1, 0x1.d77b654b82c34p-1, 0x1.d0cb58f6ec074p-1, 0x1.ca4b3055ee191p-1,
    0x1.c3f8f01c3f8f0p-1, 0x1.bdd2b899406f7p-1, 0x1.b7d6c3dda338bp-1, 0x1.b2036406c80d9p-1,
    0x1.ac5701ac5701bp-1, 0x1.a6d01a6d01a6dp-1, 0x1.a16d3f97a4b02p-1, 0x1.9c2d14ee4a102p-1,
    0x1.970e4f80cb872p-1, 0x1.920fb49d0e229p-1, 0x1.8d3018d3018d3p-1, 0x1.886e5f0abb04ap-1,
    0x1.83c977ab2beddp-1, 0x1.7f405fd017f40p-1, 0x1.7ad2208e0ecc3p-1, 0x1.767dce434a9b1p-1,
    0x1.724287f46debcp-1, 0x1.6e1f76b4337c7p-1, 0x1.6a13cd1537290p-1, 0x1.661ec6a5122f9p-1,
    0x1.623fa77016240p-1, 0x1.5e75bb8d015e7p-1, 0x1.5ac056b015ac0p-1, 0x1.571ed3c506b3ap-1,
    0x1.5390948f40febp-1, 0x1.5015015015015p-1, 0x1.4cab88725af6ep-1, 0x1.49539e3b2d067p-1,
    0x1.460cbc7f5cf9ap-1, 0x1.42d6625d51f87p-1, 0x1.3fb013fb013fbp-1, 0x1.3c995a47babe7p-1,
    0x1.3991c2c187f63p-1, 0x1.3698df3de0748p-1, 0x1.33ae45b57bcb2p-1, 0x1.30d190130d190p-1,
    0x1.2e025c04b8097p-1, 0x1.2b404ad012b40p-1, 0x1.288b01288b013p-1, 0x1.25e22708092f1p-1,
    0x1.23456789abcdfp-1, 0x1.20b470c67c0d9p-1, 0x1.1e2ef3b3fb874p-1, 0x1.1bb4a4046ed29p-1,
    0x1.19453808ca29cp-1, 0x1.16e0689427379p-1, 0x1.1485f0e0acd3bp-1, 0x1.12358e75d3033p-1,
    0x1.0fef010fef011p-1, 0x1.0db20a88f4696p-1, 0x1.0b7e6ec259dc8p-1, 0x1.0953f39010954p-1,
    0x1.073260a47f7c6p-1, 0x1.05197f7d73404p-1, 0x1.03091b51f5e1ap-1, 0x1.0101010101010p-1};
__constant__ double LOG_NEG_LOG_INV_C[64] = {
    0x1.fe02a6b106799p-8, 0x1.7b91b07d5b126p-6, 0x1.39e87b9febd68p-5, 0x1.b42dd711971b9p-5,
    0x1.16536eea37ae3p-4, 0x1.51b073f06183cp-4, 0x1.8c345d6319b23p-4, 0x1.c5e548f5bc743p-4,
    0x1.fec9131dbeabcp-4, 0x1.1b72ad52f67a2p-3, 0x1.371fc201e8f75p-3, 0x1.526e5e3a1b438p-3,
    0x1.6d60fe719d21bp-3, 0x1.87fa06520c911p-3, 0x1.a23bc1fe2b561p-3, 0x1.bc286742d8cd4p-3,
    0x1.d5c216b4fbb94p-3, 0x1.ef0adcbdc5935p-3, 0x1.0402594b4d041p-2, 0x1.1058bf9ae4ad4p-2,
    0x1.1c898c16999fbp-2, 0x1.2895a13de86a4p-2, 0x1.347dd9a987d56p-2, 0x1.404308686a7e4p-2,
    0x1.4be5f957778a1p-2, 0x1.5767717455a6cp-2, 0x1.62c82f2b9c796p-2, 0x1.6e08eaa2ba1e4p-2,
    0x1.792a55fdd47a1p-2, 0x1.842d1da1e8b18p-2, 0x1.8f11e873662c8p-2, 0x1.99d958117e08ap-2,
    0x1.a484090e5bb09p-2, 0x1.af1293247786bp-2, 0x1.b9858969310fdp-2, 0x1.c3dd7a7cdad4dp-2,
    0x1.ce1af0b85f3ecp-2, 0x1.d83e7258a2f3ep-2, 0x1.e24881a7c6c26p-2, 0x1.ec399d2468cc1p-2,
    0x1.f6123fa7028adp-2, 0x1.ffd2e0857f497p-2, 0x1.04bdf9da926d2p-1, 0x1.0986f4f573521p-1,
    0x1.0e44985d1cc8cp-1, 0x1.12f719593efbdp-1, 0x1.179eabbd899a0p-1, 0x1.1c3b81f713c25p-1,
    0x1.20cdcd192ab6ep-1, 0x1.2555bce98f7cap-1, 0x1.29d37fec2b08bp-1, 0x1.2e47436e40268p-1,
    0x1.32b1339121d71p-1, 0x1.37117b54747b6p-1, 0x1.3b68449fffc23p-1, 0x1.3fb5b84d16f43p-1,
    0x1.43f9fe2f9ce67p-1, 0x1.48353d1ea88dfp-1, 0x1.4c679afccee39p-1, 0x1.50913cc01686bp-1,
    0x1.54b2467999498p-1, 0x1.58cadb5cd7989p-1, 0x1.5cdb1dc6c1765p-1, 0x1.60e32f44788d9p-1};

// sm_log: [0, 64) = LOG_INV_C, [64, 128) = LOG_NEG_LOG_INV_C (shared memory: any index pattern, no serialisation)
__device__ __forceinline__ double log_tab64(double x, const double* __restrict__ sm_log) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const int j = (hi >> 14) & 63;
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
    const double r = fma(m, sm_log[j], -1.0);
    double q = fma(r, 0x1.2492492492492p-3, -0x1.5555555555555p-3);  // 1/7, -1/6
    q = fma(q, r, 0x1.999999999999ap-3);                              // 1/5
    q = fma(q, r, -0.25);
    q = fma(q, r, 0x1.5555555555555p-2);                              // 1/3
    q = fma(q, r, -0.5);
    const double l1p = fma(r * r, q, r);
    const double ed = (double)e;
    double lg = fma(ed, 1.90821492927058770002e-10, l1p) + sm_log[64 + j];
    return fma(ed, 6.93147180369123816490e-01, lg);  // ln2 split as in exp_core: e * hi is exact
}

__global__ void k_normal_logprob(const double* __restrict__ loc, const double* __restrict__ scale,
                                 const double* __restrict__ x, size_t n, double* __restrict__ ll) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n) return;
    double z = (x[gid] - loc[gid]) / scale[gid];
    ll[gid] = -0.5 * z * z - 0.5 * LOG_2PI - log(scale[gid]);
}

// Normal branch (weights.py:95-96: elementwise Normal log-pdf with scale = covariance, mean over the observation
// realisations, exp, normalise).  The log-density is quadratic in the observation, so its mean needs only the
// sample mean and the centred second moment of the point's Ro observations:
//   mean_r ll = -1/2 (var_o + (mean_o - loc)^2) / scale^2 - 1/2 log 2pi - log scale,   var_o = mean_r (o_r - mean_o)^2
// (centred, so nothing cancels when loc is close to the observations; the residual of the rounded mean is carried) -- one division and one log per model
// instead of Ro divisions; scales that are not ordinary numbers (zero, subnormal, huge, infinite, NaN) keep
// the per-realisation form and with it the reference's inf / NaN results.
__global__ void k_loglik_weights_normal(const double* __restrict__ loc, const double* __restrict__ scale,
                                        const double* __restrict__ obs, int C, int M, int Ro, int N, double cst,
                                        double* __restrict__ w, double* __restrict__ lls_exp,
                                        double* __restrict__ lls_mean, int smem_ok) {
    extern __shared__ double wstage[];
    __shared__ double tab[16];
    __shared__ double ltab[128];
    if (threadIdx.x < 16) tab[threadIdx.x] = EXP2_16TH[threadIdx.x];
    for (int e = threadIdx.x; e < 128; e += blockDim.x) ltab[e] = e < 64 ? LOG_INV_C[e] : LOG_NEG_LOG_INV_C[e - 64];
    __syncthreads();
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)C * N) return;
    int c = (int)(gid / N), i = (int)(gid % N);
    WeightStage st(wstage, smem_ok, w + (size_t)c * M * N + i, (size_t)N);
    const double* ob = obs + (size_t)c * Ro * N + i;
    double mean_o = 0.0;
    for (int r = 0; r < Ro; ++r) mean_o += ob[(size_t)r * N];
    mean_o /= Ro;
    double var_o = 0.0, mean_d = 0.0;  // mean_d: what the rounded pivot mean_o leaves of mean_r (o_r - mean_o)
    for (int r = 0; r < Ro; ++r) {
        const double d = ob[(size_t)r * N] - mean_o;
        var_o = fma(d, d, var_o);
        mean_d += d;
    }
    var_o /= Ro;
    mean_d /= Ro;
    const double two_mean_d = 2.0 * mean_d;
    double total = 0.0;
    const size_t o0 = (size_t)c * M * N + i;
    // the general member: library log / exp, the per-realisation form for scales that are not ordinary numbers
    auto slow_member = [&](int m) {
        const size_t o = o0 + (size_t)m * N;
        const double l = loc[o], sc = scale[o];
        const double lsc = log(sc);
        double mean;
        if (sc > 0x1p-500 && sc < 0x1p500) {
            const double dl = mean_o - l;  // (o_r - loc) = d_r + dl exactly, for any pivot
            mean = (-0.5 * (fma(dl, two_mean_d + dl, var_o) * __drcp_rn(sc * sc)) - 0.5 * LOG_2PI) - lsc;
        } else {
            double s = 0.0;
            for (int r = 0; r < Ro; ++r) {
                double z = (ob[(size_t)r * N] - l) / sc;
                s += -0.5 * z * z - 0.5 * LOG_2PI - lsc;
            }
            mean = s / Ro;
        }
        const double x = cst * mean;
        double e = exp_tab16_core(x, tab);
        if (!exp_tab16_ok(x)) e = exp(x);
        if (lls_mean) lls_mean[o] = mean;
        if (lls_exp) lls_exp[o] = e;
        st[m] = e;
        if (e == e) total += e;  // xarray .sum('model') skips NaN (weights.py:122)
    };
    // members four at a time: the loads are issued together and the four log / reciprocal / exp chains are independent
    // and branch-free (table log, rounded reciprocal, table exp) -- ~45 instead of ~130 instructions per member.
    // A group with a scale or an exponent outside the tables' range goes through the general member one by one.
    int m = 0;
    for (; m + 4 <= M; m += 4) {
        double l[4], sc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            l[j] = loc[o0 + (size_t)(m + j) * N];
            sc[j] = scale[o0 + (size_t)(m + j) * N];
        }
        bool ok = true;
        double mean[4], x[4], e[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ok = ok && sc[j] > 0x1p-500 && sc[j] < 0x1p500;
            const double lsc = log_tab64(sc[j], ltab);
            const double dl = mean_o - l[j];
            mean[j] = (-0.5 * (fma(dl, two_mean_d + dl, var_o) * __drcp_rn(sc[j] * sc[j])) - 0.5 * LOG_2PI) - lsc;
            x[j] = cst * mean[j];
            e[j] = exp_tab16_core(x[j], tab);
            ok = ok && exp_tab16_ok(x[j]);
        }
        if (ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const size_t o = o0 + (size_t)(m + j) * N;
                if (lls_mean) lls_mean[o] = mean[j];
                if (lls_exp) lls_exp[o] = e[j];
                st[m + j] = e[j];
                total += e[j];  // finite by construction (|x| < 700)
            }
        } else {
            for (int j = 0; j < 4; ++j) slow_member(m + j);
        }
    }
    for (; m < M; ++m) slow_member(m);
    if (total > 0x1p-1000 && total < 0x1p1000) {  // ordinary normaliser: one reciprocal, M products (as the MVN kernel)
        const double inv = 1.0 / total;
#pragma unroll 4
        for (int mm = 0; mm < M; ++mm) w[o0 + (size_t)mm * N] = st[mm] * inv;
    } else {  // zero, subnormal, huge, infinite, NaN: the divisions, so that 0/0, x/inf and NaN come out as in the reference
        for (int mm = 0; mm < M; ++mm) w[o0 + (size_t)mm * N] = st[mm] / total;
    }
}

// member-sharded normalisation: w = lls_exp / total  (weights.py:122-123 after an all-reduce of the sum)
__global__ void k_weights_normalise(const double* __restrict__ lls_exp, const double* __restrict__ total, int C, int M,
                                    int T, double* __restrict__ w) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)C * M * T) return;
    int i = (int)(gid % T);
    int c = (int)(gid / ((size_t)M * T));
    w[gid] = lls_exp[gid] / total[(size_t)c * T + i];
}

// xarray .mean('time') skips NaN (utils.py:111); broadcast back over time (utils.py:133)
__global__ void __launch_bounds__(256) k_weights_time_mean(const double* __restrict__ w, int CM, int T,
                                                           double* __restrict__ out) {
    __shared__ double rs[8];
    __shared__ double rc[8];
    int row = blockIdx.x;
    const double* p = w + (size_t)row * T;
    double s = 0.0, cnt = 0.0;
    for (int i = threadIdx.x; i < T; i += 256) {
        double v = p[i];
        if (!isnan(v)) {
            s += v;
            cnt += 1.0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) {
        rs[threadIdx.x >> 5] = s;
        rc[threadIdx.x >> 5] = cnt;
    }
    __syncthreads();
    double S = 0, Cn = 0;
    for (int q = 0; q < 8; ++q) {
        S += rs[q];
        Cn += rc[q];
    }
    double m = S / Cn;  // all-NaN row -> 0/0 = NaN, as xarray
    for (int i = threadIdx.x; i < T; i += 256) out[(size_t)row * T + i] = m;
}

// --------------------------------------------------------------------------------------
// a5/a6: 1-D Gaussian W2 barycentre with the reference's SIGNED stop rule (wasserstein.py:88)
// --------------------------------------------------------------------------------------
__device__ __forceinline__ void barycentre_iterate(double S_unit /* sum w s, or NaN */, bool use_unit,
                                                   const double* __restrict__ w, const double* __restrict__ var,
                                                   size_t stride, int M, double tol, double init_var, int max_iters,
                                                   double& sigma, int& iters, bool have_first = false,
                                                   double first_cand = 0.0) {
    double bv = init_var;
    int n_it = 0;
    while (true) {
        double cand = 0.0;
        double sq = sqrt(bv);
        if (use_unit) {
            cand = sq * S_unit;
        } else if (n_it == 0 && have_first) {
            cand = first_cand;  // formed by the caller in the pass that also read the means
        } else {
            for (int m = 0; m < M; ++m) cand += w[m * stride] * sq * sqrt(var[m * stride]);
        }
        if (cand - bv < tol) {
            bv = cand;
            break;
        }
        if (cand != cand) {  // NaN never satisfies the signed test and never leaves: the loop would only count to
            bv = cand;       // max_iters + 1 ("not converged", the reference warns) -- same outputs, no spinning
            n_it = max_iters + 1;
            break;
        }
        bv = cand;
        ++n_it;
        if (n_it > max_iters) break;
    }
    sigma = sqrt(bv);
    iters = n_it;
}

__global__ void k_barycentre_1d(const double* __restrict__ means, const double* __restrict__ variances,
                                const double* __restrict__ weights, int C, int M, int N, double tol, double init_var,
                                int max_iters, double* __restrict__ mu, double* __restrict__ sigma,
                                int* __restrict__ iters) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)C * N) return;
    int c = (int)(gid / N), i = (int)(gid % N);
    size_t base = (size_t)c * M * N + i;
    // one pass over the three arrays: the weighted mean (wasserstein.py:98) and the first candidate of the
    // variance iteration (:85-86), which is also the last one whenever sum w sigma < init_var (quirk Q-BARY)
    double m_acc = 0.0, cand0 = 0.0;
    const double sq0 = sqrt(init_var);
#pragma unroll 4
    for (int m = 0; m < M; ++m) {
        const double wm = weights[base + (size_t)m * N];
        m_acc += wm * means[base + (size_t)m * N];
        cand0 += wm * sq0 * sqrt(variances[base + (size_t)m * N]);
    }
    double sg;
    int it;
    barycentre_iterate(0.0, false, weights + base, variances + base, (size_t)N, M, tol, init_var, max_iters, sg, it, true,
                       cand0);
    mu[gid] = m_acc;
    sigma[gid] = sg;
    if (iters) iters[gid] = it;
}

__global__ void k_barycentre_partial(const double* __restrict__ means, const double* __restrict__ variances,
                                     const double* __restrict__ lls_exp, int C, int M, int N,
                                     double* __restrict__ partial) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t CN = (size_t)C * N;
    if (gid >= CN) return;
    int c = (int)(gid / N), i = (int)(gid % N);
    size_t base = (size_t)c * M * N + i;
    double s0 = 0, s1 = 0, s2 = 0;
    for (int m = 0; m < M; ++m) {
        double w = lls_exp[base + (size_t)m * N];
        if (w == w) s0 += w;  // the normaliser skips NaN members (xarray .sum('model')); the barycentre sums do not
        s1 += w * means[base + (size_t)m * N];
        s2 += w * sqrt(variances[base + (size_t)m * N]);
    }
    partial[gid] = s0;
    partial[CN + gid] = s1;
    partial[2 * CN + gid] = s2;
}

__global__ void k_barycentre_finish(const double* __restrict__ partial, int C, int N, double tol, double init_var,
                                    int max_iters, double* __restrict__ mu, double* __restrict__ sigma,
                                    int* __restrict__ iters) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t CN = (size_t)C * N;
    if (gid >= CN) return;
    double s0 = partial[gid];
    double S = partial[2 * CN + gid] / s0;
    double sg;
    int it;
    barycentre_iterate(S, true, nullptr, nullptr, 0, 0, tol, init_var, max_iters, sg, it);
    mu[gid] = partial[CN + gid] / s0;
    sigma[gid] = sg;
    if (iters) iters[gid] = it;
}

}  // namespace be
