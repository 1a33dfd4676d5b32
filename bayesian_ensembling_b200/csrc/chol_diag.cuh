// Diagonal-block kernel of the blocked Cholesky / triangular inverse (sm_100a, fp64).
//
// One CTA (8 warps) factors one (<=128)^2 diagonal block in shared memory and inverts the
// factor; this is the latency-bound serial spine of the factorisation, so it is organised to
// keep the dependent chain short (an earlier version that factored 32 x 32 sub-blocks in one
// warp with shuffles spent 480k cycles per block, all of it one warp's instruction stream):
//   * 8-column steps.  Thread t owns row t.  Every row-owning thread REDUNDANTLY factors the
//     8 x 8 pivot block in its own registers (36 values, 8 rsqrt, broadcast loads) and then
//     eliminates its own row with it -- no shuffles, no inter-warp hand-off, one barrier.
//   * The trailing columns are updated by ALL warps with FP64 tensor-core DMMA m8n8k4
//     fragments read straight from shared memory (row stride 132 doubles == 4 mod 16 makes
//     both the row-major and the transposed fragment loads bank-conflict free).
//   * The 128 x 128 inverse is built in place by recursive doubling (8 -> 16 -> ... -> 128),
//     inv([[A,0],[B,C]]) = [[A^-1,0],[-C^-1 B A^-1, C^-1]], each level two DMMA products.
//
// Padding semantics are those of be_kernels.cuh: columns >= T never pivot; rows >= T inside the
// block (right-hand sides riding along) are treated as extra panel rows.
#pragma once
#include "dmma_gemm.cuh"

namespace be {

constexpr int DG_LD = 132;
constexpr int DG_W = 8;        // columns per elimination step
constexpr int DG_TMP_LD = 68;  // == 4 mod 16
constexpr int DG_SMEM_DOUBLES = 128 * DG_LD + 64 * DG_TMP_LD;
constexpr int DG_SMEM_BYTES = DG_SMEM_DOUBLES * 8;

// out[f] += sum_k A[g][k] * B[f*8 + g][k]   ("NT": both operands k-contiguous rows)
template <int NF, int K>
__device__ __forceinline__ void frag_nt(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                        double (&acc)[NF][2]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int k = 0; k < K; k += 4) {
        double a = A[g * lda + k + q];
#pragma unroll
        for (int f = 0; f < NF; ++f) dmma884(acc[f][0], acc[f][1], a, B[(f * 8 + g) * ldb + k + q]);
    }
}
// out[f] += sum_k A[g][k] * B[k][f*8 + g]   ("NN")
template <int NF>
__device__ __forceinline__ void frag_nn(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                        int K, double (&acc)[NF][2]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
#pragma unroll 4
    for (int k = 0; k < K; k += 4) {
        double a = A[g * lda + k + q];
#pragma unroll
        for (int f = 0; f < NF; ++f) dmma884(acc[f][0], acc[f][1], a, B[(k + q) * ldb + f * 8 + g]);
    }
}

// 1/sqrt(x) to ~1 ulp without the library's special-case branches: MUFU seed + two Newton
// steps (x <= 0 gives NaN/inf, which propagates like a failed LAPACK pivot would).
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        double e = fma(-x * y, y, 1.0);
        y = fma(0.5 * y, e, y);
    }
    return y;
}

// One recursive-doubling product over all pairs of a level: for pair p (blocks of size s)
//   OUT_p[s x s] = sign * X_p[s x s] * Y_p[s x s]        (row-major, "NN")
// X_p = X + p * xs, etc.  Work is split in 8 x 16 output pieces over the 8 warps.
__device__ __forceinline__ void level_product(const double* X, int ldx, int xs, const double* Y, int ldy, int ys,
                                              double* OUT, int ldo, int os, int s, int npairs, double sign) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int fr_n = s / 8, fc_n = (s + 15) / 16;
    const int per_pair = fr_n * fc_n;
    for (int task = warp; task < npairs * per_pair; task += 8) {
        int p = task / per_pair, rem = task % per_pair;
        int fr = rem / fc_n, fc = rem % fc_n;
        const double* A = X + (size_t)p * xs + fr * 8 * ldx;
        const double* Bm = Y + (size_t)p * ys + fc * 16;
        double* O = OUT + (size_t)p * os + (fr * 8 + g) * ldo + fc * 16 + 2 * q;
        if (s >= 16) {
            double acc[2][2] = {};
            frag_nn<2>(A, ldx, Bm, ldy, s, acc);
            *reinterpret_cast<double2*>(O) = make_double2(sign * acc[0][0], sign * acc[0][1]);
            *reinterpret_cast<double2*>(O + 8) = make_double2(sign * acc[1][0], sign * acc[1][1]);
        } else {
            double acc[1][2] = {};
            frag_nn<1>(A, ldx, Bm, ldy, s, acc);
            *reinterpret_cast<double2*>(O) = make_double2(sign * acc[0][0], sign * acc[0][1]);
        }
    }
}

// Mat block kb (rows/cols r0 .. r0+n) -> Cholesky factor in place (lower, real columns only),
// Dinv[b][kb] = blockdiag(L11^-1, I) (row-major 128 x 128) and optionally V tile (kb,kb) = Dinv^T.
__global__ void __launch_bounds__(256, 1) k_diag_block(double* __restrict__ Mat, int ld, int Tp, int T, int kb,
                                                       double* __restrict__ Dinv, int nblk, double* __restrict__ V,
                                                       int* __restrict__ info) {
    extern __shared__ double sm[];
    double* S = sm;                  // [128][132]
    double* TMP = sm + 128 * DG_LD;  // [64][68]
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const int r0 = kb * 128;
    const int n = min(128, Tp - r0);        // rows present (multiple of 16)
    const int nr = max(0, min(n, T - r0));  // real (pivoting) columns
    double* Mb = Mat + (size_t)b * Tp * ld;
    // load the block (lower part, real columns); 8 independent global loads in flight per thread
#pragma unroll 1
    for (int e0 = tid; e0 < 128 * 128; e0 += 256 * 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            int e = e0 + u * 256, i = e >> 7, j = e & 127;
            v[u] = (i < n && j < nr && j <= i) ? Mb[(size_t)(r0 + i) * ld + r0 + j] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            int e = e0 + u * 256;
            S[(e >> 7) * DG_LD + (e & 127)] = v[u];
        }
    }
    __shared__ int s_bad;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    const int nsteps = (nr + DG_W - 1) / DG_W;
    int bad = 0;
#pragma unroll 1
    for (int st = 0; st < nsteps; ++st) {
        const int c0 = st * DG_W;
        const int w = min(DG_W, nr - c0);
        // ---- phase A: every row-owning thread factors the 8 x 8 pivot block redundantly in its
        // registers and eliminates its own row with it.  Branch-free: pivot rows >= w (padding or
        // right-hand-side rows inside the band) enter the local copy as identity rows, so padded
        // columns are never touched; the pivot rows themselves are produced by the same
        // elimination with the columns right of their diagonal masked at the store.
        const bool act = tid < 128 && tid >= c0 && tid < n;
        double Lb[DG_W][DG_W];
        double pv[DG_W];
        if (act) {
#pragma unroll
            for (int i = 0; i < DG_W; ++i)
#pragma unroll
                for (int k = 0; k <= i; ++k) {
                    double v = S[(c0 + i) * DG_LD + c0 + k];
                    Lb[i][k] = (i < w) ? v : (i == k ? 1.0 : 0.0);
                }
            const double* row = S + tid * DG_LD + c0;
#pragma unroll
            for (int k = 0; k < DG_W; k += 2) {
                double2 t2 = *reinterpret_cast<const double2*>(row + k);
                pv[k] = t2.x;
                pv[k + 1] = t2.y;
            }
        }
        __syncthreads();  // every thread holds its copy of the pivot block before rows are rewritten
        if (act) {
#pragma unroll
            for (int j = 0; j < DG_W; ++j) {
                const double piv = Lb[j][j];
                bad = (bad == 0 && !(piv > 0.0)) ? r0 + c0 + j + 1 : bad;
                const double rs = fast_rsqrt(piv);
#pragma unroll
                for (int i = j + 1; i < DG_W; ++i) Lb[i][j] *= rs;
#pragma unroll
                for (int k = j + 1; k < DG_W; ++k)
#pragma unroll
                    for (int i = k; i < DG_W; ++i) Lb[i][k] = fma(-Lb[i][j], Lb[k][j], Lb[i][k]);
                pv[j] *= rs;
#pragma unroll
                for (int k = j + 1; k < DG_W; ++k) pv[k] = fma(-pv[j], Lb[k][j], pv[k]);
            }
            if (tid == c0 && bad != 0 && s_bad == 0) s_bad = bad;
            double* row = S + tid * DG_LD + c0;
#pragma unroll
            for (int k = 0; k < DG_W; k += 2) {
                double2 o;
                o.x = (c0 + k <= tid) ? pv[k] : 0.0;
                o.y = (c0 + k + 1 <= tid) ? pv[k + 1] : 0.0;
                *reinterpret_cast<double2*>(row + k) = o;
            }
        }
        __syncthreads();
        // ---- phase B: trailing real columns [p0, nr): S[r][c] -= P[r][0:8] . P[c][0:8], r >= c
        const int p0 = c0 + DG_W;
        const int ncs = (nr - p0 + 7) / 8;  // column strips (<= 0: nothing left)
        const int nrs = (n - p0) / 8;       // row strips
        if (ncs > 0) {
#pragma unroll 1
            for (int fr = warp; fr < nrs; fr += 8) {
                const double* Pr = S + (p0 + fr * 8) * DG_LD + c0;
                const double a0 = Pr[g * DG_LD + q], a1 = Pr[g * DG_LD + 4 + q];
                const int fc_end = min(fr, ncs - 1);
                const int r = p0 + fr * 8 + g;
#pragma unroll 1
                for (int fc0 = 0; fc0 <= fc_end; fc0 += 4) {
                    double b0[4], b1[4];
                    double2 cv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {  // loads first: 4 independent fragments in flight
                        const int fc = min(fc0 + u, fc_end);
                        const double* Pc = S + (p0 + fc * 8) * DG_LD + c0;
                        b0[u] = Pc[g * DG_LD + q];
                        b1[u] = Pc[g * DG_LD + 4 + q];
                        cv[u] = *reinterpret_cast<const double2*>(S + r * DG_LD + p0 + fc * 8 + 2 * q);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        double c[2] = {0.0, 0.0};
                        dmma884(c[0], c[1], a0, b0[u]);
                        dmma884(c[0], c[1], a1, b1[u]);
                        const int fc = fc0 + u;
                        const int cc = p0 + fc * 8 + 2 * q;
                        if (fc <= fc_end) {
                            double2 o;
                            o.x = (cc < nr && cc <= r) ? cv[u].x - c[0] : cv[u].x;
                            o.y = (cc + 1 < nr && cc + 1 <= r) ? cv[u].y - c[1] : cv[u].y;
                            *reinterpret_cast<double2*>(S + r * DG_LD + cc) = o;
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0 && s_bad != 0 && info) {
        if (info[b] == 0) info[b] = s_bad;
    }
    // the factor (lower, all n rows, real columns only)
    for (int e = tid; e < 128 * 128; e += 256) {
        int r = e >> 7, c = e & 127;
        if (r < n && c <= r && c < nr) Mb[(size_t)(r0 + r) * ld + r0 + c] = S[r * DG_LD + c];
        if (r < nr && c > r && c < n) Mb[(size_t)(r0 + r) * ld + r0 + c] = 0.0;  // clean strict upper part
    }
    __syncthreads();
    // ---- inverse of blockdiag(L11, I): rows >= nr become identity rows first
    for (int e = tid; e < 128 * 128; e += 256) {
        int r = e >> 7, c = e & 127;
        if (r >= nr && c <= r) S[r * DG_LD + c] = r == c ? 1.0 : 0.0;
    }
    __syncthreads();
    // level 0: the sixteen 8 x 8 diagonal blocks; thread t < 128 owns column (t & 7) of block (t >> 3)
    {
        double x[DG_W];
        const int blk = tid >> 3, cidx = tid & 7;
        if (tid < 128) {
            const double* Lb = S + (blk * 8) * DG_LD + blk * 8;
#pragma unroll
            for (int i = 0; i < DG_W; ++i) {
                double sacc = (i == cidx) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < i; ++k) sacc = fma(-Lb[i * DG_LD + k], (k >= cidx) ? x[k] : 0.0, sacc);
                x[i] = (i >= cidx) ? sacc / Lb[i * DG_LD + i] : 0.0;
            }
        }
        __syncthreads();
        if (tid < 128) {
            double* Lb = S + (blk * 8) * DG_LD + blk * 8;
#pragma unroll
            for (int i = 0; i < DG_W; ++i)
                if (i >= cidx) Lb[i * DG_LD + cidx] = x[i];
        }
        __syncthreads();
    }
    // levels s = 8, 16, 32, 64: B <- -C^-1 * (B * A^-1) for every pair [[A,0],[B,C]] of size 2s
    for (int s = 8; s <= 64; s *= 2) {
        const int npairs = 64 / s;
        const int stride = 2 * s * DG_LD + 2 * s;  // from one pair's A to the next
        // T_p = B_p * Ainv_p  -> TMP (pair p at column offset p * s ... rows 0..s)
        level_product(S + s * DG_LD, DG_LD, stride, S, DG_LD, stride, TMP, DG_TMP_LD, s, s, npairs, 1.0);
        __syncthreads();
        // B_p = -Cinv_p * T_p
        level_product(S + s * DG_LD + s, DG_LD, stride, TMP, DG_TMP_LD, s, S + s * DG_LD, DG_LD, stride, s, npairs, -1.0);
        __syncthreads();
    }
    double* Db = Dinv + ((size_t)b * nblk + kb) * 128 * 128;
    for (int e = tid; e < 128 * 128; e += 256) {
        int r = e >> 7, c = e & 127;
        Db[e] = c <= r ? S[r * DG_LD + c] : 0.0;
    }
    if (V) {
        double* Vb = V + (size_t)b * Tp * ld;
        for (int e = tid; e < 128 * 128; e += 256) {
            int r = e >> 7, c = e & 127;  // V tile entry (r, c) = Dinv[c][r]
            if (r >= n || c >= n) continue;
            Vb[(size_t)(r0 + r) * ld + r0 + c] = r <= c ? S[c * DG_LD + r] : 0.0;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// k_diag_block2: the same algorithm in 100 KB of shared memory instead of 170 KB, so that TWO CTAs share an SM.
// The one-CTA-per-SM version is latency-bound (ncu r01g: issue active 28 %, 12 % of the warp slots): nothing
// covers its barrier waits, its dependent pivot chains or the latency of its global loads and stores.
//   * Only the lower trapezoid is stored: rows 0..63 keep 64 columns (row stride 68), rows 64..127 keep 128
//     (row stride 132); both strides are 4 mod 16, so every fragment load stays bank-conflict free.
//   * The recursive-doubling inverse works IN PLACE, without the 64 x 68 scratch tile: T = B A^-1 is formed
//     one 8-row strip per warp (the whole strip is accumulated in registers before it overwrites B's strip),
//     and B <- -C^-1 T one 8-column block per warp (the block's operand fragments are loaded into registers
//     before its first tile is overwritten).  Every level has exactly eight strips / blocks: one per warp.
// ------------------------------------------------------------------------------------------------
constexpr int DG2_LDT = 68;                  // rows 0..63
constexpr int DG2_LDB = 132;                 // rows 64..127
constexpr int DG2_TOP = 64 * DG2_LDT;
constexpr int DG2_SMEM_DOUBLES = DG2_TOP + 64 * DG2_LDB;
constexpr int DG2_SMEM_BYTES = DG2_SMEM_DOUBLES * 8;

__device__ __forceinline__ int dg2_ld(int r) { return r < 64 ? DG2_LDT : DG2_LDB; }
__device__ __forceinline__ int dg2_off(int r, int c) {
    return r < 64 ? r * DG2_LDT + c : DG2_TOP + (r - 64) * DG2_LDB + c;
}

// T strip (8 rows x s columns) = B strip * Ainv, accumulated completely, then written over the B strip.
template <int S_>
__device__ __forceinline__ void dg2_strip_product(double* Bs, int ldb, const double* Ainv, int lda) {
    constexpr int NF = S_ / 8;
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    double acc[NF][2];
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f][0] = acc[f][1] = 0.0;
#pragma unroll 4
    for (int k = 0; k < S_; k += 4) {
        const double a = Bs[g * ldb + k + q];
#pragma unroll
        for (int f = 0; f < NF; ++f) dmma884(acc[f][0], acc[f][1], a, Ainv[(k + q) * lda + f * 8 + g]);
    }
    __syncwarp();
#pragma unroll
    for (int f = 0; f < NF; ++f)
        *reinterpret_cast<double2*>(Bs + g * ldb + f * 8 + 2 * q) = make_double2(acc[f][0], acc[f][1]);
}

// B block (s rows x 8 columns) <- -Cinv * (T block): T's fragments first, then tile by tile in place.
template <int S_>
__device__ __forceinline__ void dg2_block_product(double* Tb, int ldt, const double* Cinv, int ldc) {
    constexpr int NK = S_ / 4;
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    double bf[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) bf[k] = Tb[(4 * k + q) * ldt + g];
    __syncwarp();
#pragma unroll 2
    for (int fr = 0; fr < S_ / 8; ++fr) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int k = 0; k < NK; ++k) dmma884(c0, c1, Cinv[(fr * 8 + g) * ldc + 4 * k + q], bf[k]);
        *reinterpret_cast<double2*>(Tb + (fr * 8 + g) * ldt + 2 * q) = make_double2(-c0, -c1);
    }
}

template <int S_>
__device__ __forceinline__ void dg2_level(double* S) {
    // pairs [[A,0],[B,C]] of size 2 S_ along the diagonal; eight tasks per product, one per warp
    const int warp = threadIdx.x >> 5;
    constexpr int PER = S_ / 8;            // strips (or blocks) per pair
    const int p = warp / PER, sub = warp % PER;
    const int r0 = p * 2 * S_;             // first row / column of the pair
    const int ld_b = dg2_ld(r0 + S_);      // B and C rows live in one region (2 S_ divides 64 or the pair is the whole block)
    double* Bp = S + dg2_off(r0 + S_, r0);
    const double* Ainv = S + dg2_off(r0, r0);
    const double* Cinv = S + dg2_off(r0 + S_, r0 + S_);
    dg2_strip_product<S_>(Bp + sub * 8 * ld_b, ld_b, Ainv, dg2_ld(r0));
    __syncthreads();
    dg2_block_product<S_>(Bp + sub * 8, ld_b, Cinv, ld_b);
    __syncthreads();
}

__global__ void __launch_bounds__(256, 2) k_diag_block2(double* __restrict__ Mat, int ld, int Tp, int T, int kb,
                                                        double* __restrict__ Dinv, int nblk, double* __restrict__ V,
                                                        int* __restrict__ info) {
    extern __shared__ double S[];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const int r0 = kb * 128;
    const int n = min(128, Tp - r0);        // rows present (multiple of 16)
    const int nr = max(0, min(n, T - r0));  // real (pivoting) columns
    double* Mb = Mat + (size_t)b * Tp * ld;
    // load the stored trapezoid (lower part, real columns; zero elsewhere): a warp per row, four columns per lane
#pragma unroll 4
    for (int rr = 0; rr < 16; ++rr) {
        const int i = warp + 8 * rr, j = lane * 4;
        double2 lo = make_double2(0.0, 0.0), hi = lo;
        if (i < n && j < n && j <= i) {  // n is a multiple of 16: the quad is inside the row
            const double2* src = reinterpret_cast<const double2*>(Mb + (size_t)(r0 + i) * ld + r0 + j);
            lo = src[0];
            hi = src[1];
        }
        lo.x = (j < nr && j <= i) ? lo.x : 0.0;
        lo.y = (j + 1 < nr && j + 1 <= i) ? lo.y : 0.0;
        hi.x = (j + 2 < nr && j + 2 <= i) ? hi.x : 0.0;
        hi.y = (j + 3 < nr && j + 3 <= i) ? hi.y : 0.0;
        if (i >= 64 || j < 64) {
            double2* dst = reinterpret_cast<double2*>(S + dg2_off(i, j));
            dst[0] = lo;
            dst[1] = hi;
        }
    }
    __shared__ int s_bad;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    const int nsteps = (nr + DG_W - 1) / DG_W;
    int bad = 0;
#pragma unroll 1
    for (int st = 0; st < nsteps; ++st) {
        const int c0 = st * DG_W;
        const int w = min(DG_W, nr - c0);
        const bool act = tid < 128 && tid >= c0 && tid < n;
        double Lb[DG_W][DG_W];
        double pv[DG_W];
        if (act) {
            const double* piv = S + dg2_off(c0, c0);
            const int ldp = dg2_ld(c0);
#pragma unroll
            for (int i = 0; i < DG_W; ++i)
#pragma unroll
                for (int k = 0; k <= i; ++k) {
                    double v = piv[i * ldp + k];
                    Lb[i][k] = (i < w) ? v : (i == k ? 1.0 : 0.0);
                }
            const double* row = S + dg2_off(tid, c0);
#pragma unroll
            for (int k = 0; k < DG_W; k += 2) {
                double2 t2 = *reinterpret_cast<const double2*>(row + k);
                pv[k] = t2.x;
                pv[k + 1] = t2.y;
            }
        }
        __syncthreads();
        if (act) {
#pragma unroll
            for (int j = 0; j < DG_W; ++j) {
                const double piv = Lb[j][j];
                bad = (bad == 0 && !(piv > 0.0)) ? r0 + c0 + j + 1 : bad;
                const double rs = fast_rsqrt(piv);
#pragma unroll
                for (int i = j + 1; i < DG_W; ++i) Lb[i][j] *= rs;
#pragma unroll
                for (int k = j + 1; k < DG_W; ++k)
#pragma unroll
                    for (int i = k; i < DG_W; ++i) Lb[i][k] = fma(-Lb[i][j], Lb[k][j], Lb[i][k]);
                pv[j] *= rs;
#pragma unroll
                for (int k = j + 1; k < DG_W; ++k) pv[k] = fma(-pv[j], Lb[k][j], pv[k]);
            }
            if (tid == c0 && bad != 0 && s_bad == 0) s_bad = bad;
            double* row = S + dg2_off(tid, c0);
#pragma unroll
            for (int k = 0; k < DG_W; k += 2) {
                double2 o;
                o.x = (c0 + k <= tid) ? pv[k] : 0.0;
                o.y = (c0 + k + 1 <= tid) ? pv[k + 1] : 0.0;
                *reinterpret_cast<double2*>(row + k) = o;
            }
        }
        __syncthreads();
        // trailing real columns [p0, nr): S[r][c] -= P[r][0:8] . P[c][0:8], r >= c
        const int p0 = c0 + DG_W;
        const int ncs = (nr - p0 + 7) / 8;
        const int nrs = (n - p0) / 8;
        if (ncs > 0) {
#pragma unroll 1
            for (int fr = warp; fr < nrs; fr += 8) {
                const int rr = p0 + fr * 8;      // first row of the strip (strips never straddle row 64)
                const int ldr = dg2_ld(rr);
                const double* Pr = S + dg2_off(rr, c0);
                const double a0 = Pr[g * ldr + q], a1 = Pr[g * ldr + 4 + q];
                const int fc_end = min(fr, ncs - 1);
                const int r = rr + g;
                double* Crow = S + dg2_off(r, 0);
#pragma unroll 1
                for (int fc0 = 0; fc0 <= fc_end; fc0 += 4) {
                    double b0[4], b1[4];
                    double2 cv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int fc = min(fc0 + u, fc_end);
                        const int rc = p0 + fc * 8;
                        const int ldc = dg2_ld(rc);
                        const double* Pc = S + dg2_off(rc, c0);
                        b0[u] = Pc[g * ldc + q];
                        b1[u] = Pc[g * ldc + 4 + q];
                        cv[u] = *reinterpret_cast<const double2*>(Crow + p0 + fc * 8 + 2 * q);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        double c[2] = {0.0, 0.0};
                        dmma884(c[0], c[1], a0, b0[u]);
                        dmma884(c[0], c[1], a1, b1[u]);
                        const int fc = fc0 + u;
                        const int cc = p0 + fc * 8 + 2 * q;
                        if (fc <= fc_end) {
                            double2 o;
                            o.x = (cc < nr && cc <= r) ? cv[u].x - c[0] : cv[u].x;
                            o.y = (cc + 1 < nr && cc + 1 <= r) ? cv[u].y - c[1] : cv[u].y;
                            *reinterpret_cast<double2*>(Crow + cc) = o;
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0 && s_bad != 0 && info) {
        if (info[b] == 0) info[b] = s_bad;
    }
    // the factor (lower, all n rows, real columns only; the strict upper part of the real rows is cleaned)
#pragma unroll 2
    for (int rr = 0; rr < 16; ++rr) {
        const int r = warp + 8 * rr, c = lane * 4;
        if (r < n && c < n) {
            double* dst = Mb + (size_t)(r0 + r) * ld + r0 + c;
            if (c + 3 <= r && c + 3 < nr) {
                const double2* src = reinterpret_cast<const double2*>(S + dg2_off(r, c));
                reinterpret_cast<double2*>(dst)[0] = src[0];
                reinterpret_cast<double2*>(dst)[1] = src[1];
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int cc = c + u;
                    if (cc <= r && cc < nr) dst[u] = S[dg2_off(r, cc)];
                    else if (r < nr && cc > r) dst[u] = 0.0;
                }
            }
        }
    }
    __syncthreads();
    // inverse of blockdiag(L11, I): rows >= nr become identity rows first
    for (int r = nr + warp; r < 128; r += 8)
        for (int c = lane; c <= r; c += 32) S[dg2_off(r, c)] = r == c ? 1.0 : 0.0;
    __syncthreads();
    {   // level 0: the sixteen 8 x 8 diagonal blocks; thread t < 128 owns column (t & 7) of block (t >> 3)
        double x[DG_W];
        const int blk = tid >> 3, cidx = tid & 7;
        const int ldl = dg2_ld((blk & 15) * 8);
        double* Lb = S + dg2_off((blk & 15) * 8, (blk & 15) * 8);
        if (tid < 128) {
#pragma unroll
            for (int i = 0; i < DG_W; ++i) {
                double sacc = (i == cidx) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < i; ++k) sacc = fma(-Lb[i * ldl + k], (k >= cidx) ? x[k] : 0.0, sacc);
                x[i] = (i >= cidx) ? sacc / Lb[i * ldl + i] : 0.0;
            }
        }
        __syncthreads();
        if (tid < 128) {
#pragma unroll
            for (int i = 0; i < DG_W; ++i)
                if (i >= cidx) Lb[i * ldl + cidx] = x[i];
        }
        __syncthreads();
    }
    dg2_level<8>(S);
    dg2_level<16>(S);
    dg2_level<32>(S);
    dg2_level<64>(S);
    double* Db = Dinv + ((size_t)b * nblk + kb) * 128 * 128;
    double* Vb = V ? V + (size_t)b * Tp * ld : nullptr;
#pragma unroll 2
    for (int rr = 0; rr < 16; ++rr) {
        const int r = warp + 8 * rr, c = lane * 4;
        double2 lo, hi;
        lo.x = c <= r ? S[dg2_off(r, c)] : 0.0;
        lo.y = c + 1 <= r ? S[dg2_off(r, c + 1)] : 0.0;
        hi.x = c + 2 <= r ? S[dg2_off(r, c + 2)] : 0.0;
        hi.y = c + 3 <= r ? S[dg2_off(r, c + 3)] : 0.0;
        double2* dst = reinterpret_cast<double2*>(Db + r * 128 + c);
        dst[0] = lo;
        dst[1] = hi;
        if (Vb && r < n && c < n) {  // V tile entry (r, c) = Dinv[c][r], upper triangular
            double2 vlo, vhi;
            vlo.x = r <= c ? S[dg2_off(c, r)] : 0.0;
            vlo.y = r <= c + 1 ? S[dg2_off(c + 1, r)] : 0.0;
            vhi.x = r <= c + 2 ? S[dg2_off(c + 2, r)] : 0.0;
            vhi.y = r <= c + 3 ? S[dg2_off(c + 3, r)] : 0.0;
            double2* vd = reinterpret_cast<double2*>(Vb + (size_t)(r0 + r) * ld + r0 + c);
            vd[0] = vlo;
            vd[1] = vhi;
        }
    }
}

}  // namespace be
