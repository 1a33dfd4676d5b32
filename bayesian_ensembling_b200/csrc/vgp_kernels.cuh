// Kernels of the VGP training loop GPDTW1D.fit runs (ensembles/models.py:185-220): whitened
// variational GP, heteroskedastic Gaussian likelihood, natural-gradient steps on q and Adam
// steps on the two Matern-3/2 hyper-parameters, then predict_f(full_cov=True).
//
// Everything O(T^3) goes through the same FP64 DMMA tile engine as the fixed-theta path: one
// generic batched "NT" tile kernel (k_gemm_nt) whose tile set, contraction range and epilogue
// are template / argument choices, plus the blocked potrf / trtri of be_api.cu.  State is kept
// in NATURAL parameters across iterations (P = S^-1 = -2 theta_2 and n1 = theta_1), which is
// the fixed point form of gpflow's natgrad step for this conjugate model ("L2").
#pragma once
#include "be_kernels.cuh"

namespace be {

enum { SHAPE_FULL = 0, SHAPE_LOWER = 1, SHAPE_UPPER = 2 };  // block pairs (tA, tB): all / tA>=tB / tA<=tB
enum { KLO_ZERO = 0, KLO_TA = 1, KLO_MAX = 2 };             // first contracted column: 0 / tA*128 / max(tA,tB)*128
enum { KHI_END = 0, KHI_TB = 1, KHI_TA = 2 };               // one past the last: Tp / (tB+1)*128 / (tA+1)*128

struct GemmArgs {
    const double* A;
    const double* Bm;
    int lda, ldb;
    size_t strideA, strideB;  // per-problem strides (doubles)
    int divA, divB;           // problem b reads operand slot b / div (one S^1/2 shared by the M members of a cell)
    int Tp, nblk, B;
    int shape, klo, khi;
    int T;                    // real dimension (host-side flop booking only)
};

__host__ __device__ inline int gemm_tiles(int nblk, int shape) {
    return shape == SHAPE_FULL ? nblk * nblk : nblk * (nblk + 1) / 2;
}

// C-tile(tA, tB)[i, j] = sum_{k in [k0, k1)} A[tA*128 + i, k] * Bm[tB*128 + j, k]; the epilogue
// object decides what to do with every pair of adjacent columns.
template <class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, GEMM_CTAS_PER_SM) k_gemm_nt(GemmArgs g, Epi epi) {
    extern __shared__ __align__(16) double2 smem2[];
    int tile, half, b;
    cta_decode(g.B, tile, half, b);
    int tA, tB;
    if (g.shape == SHAPE_FULL) {
        tA = tile / g.nblk;
        tB = tile % g.nblk;
    } else if (g.shape == SHAPE_LOWER) {
        tri_decode(tile, tA, tB);
    } else {
        tri_decode(tile, tB, tA);
    }
    const int a_rows = blk_rows(g.Tp, tA), b_rows = min(BN, blk_rows(g.Tp, tB) - half * BN);
    if (b_rows <= 0) return;
    int k0 = g.klo == KLO_ZERO ? 0 : (g.klo == KLO_TA ? tA : max(tA, tB)) * NB;
    int k1 = g.khi == KHI_END ? g.Tp : min(g.Tp, ((g.khi == KHI_TB ? tB : tA) + 1) * NB);
    TileAcc acc;
    const int klen = max(0, k1 - k0);
    gemm_nt_mainloop(g.A + (size_t)(b / g.divA) * g.strideA + (size_t)tA * NB * g.lda + k0, g.lda, a_rows,
                     g.Bm + (size_t)(b / g.divB) * g.strideB + (size_t)(tB * NB + half * BN) * g.ldb + k0, g.ldb, b_rows, klen, smem2,
                     acc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
        int r = acc_row(warp, lane, mi);
        if (r >= a_rows) continue;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            int c = acc_col(warp, lane, ni);
            if (c >= b_rows) continue;
            epi(b, tA * NB + r, tB * NB + half * BN + c, acc.v[mi][ni][0], acc.v[mi][ni][1]);
        }
    }
    epi.finish(b, tile * 2 + half, reinterpret_cast<double*>(smem2));
}

struct EpiBase {
    __device__ void finish(int, int, double*) {}
};

// out[gr, gc] = value on the real T x T part, `pad_diag` on the padded diagonal, 0 elsewhere.
// mirror: also write the transposed entry (symmetric results computed on lower tiles only).
// sub: subtract Sub[gr, gc] first (W = AT * S - AT).
struct EpiStore : EpiBase {
    double* out;
    const double* sub;
    int ld, Tp, T;
    double pad_diag;
    int mirror;
    __device__ void operator()(int b, int gr, int gc, double v0, double v1) {
        double* ob = out + (size_t)b * Tp * ld;
        double v[2] = {v0, v1};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int c = gc + e;
            double val;
            if (gr < T && c < T) {
                val = v[e];
                if (sub) val -= sub[(size_t)b * Tp * ld + (size_t)gr * ld + c];
            } else {
                val = gr == c ? pad_diag : 0.0;
            }
            ob[(size_t)gr * ld + c] = val;
            if (mirror) ob[(size_t)c * ld + gr] = val;
        }
    }
};

// natural-gradient step on theta_2 (models.py:209):  P <- (1-gamma) P + gamma (I + L^T D^-1 L),
// lower tiles; also copied to `work`, the buffer the next factorisation destroys.
// If G != nullptr the accumulators G = L^T D^-1 L themselves are kept as a full symmetric padded matrix (zero padding)
// for the Adam half's G S product: written from the positions at or below the diagonal only, with their mirror --
// (L_ki / s_k) L_kj and (L_kj / s_k) L_ki round differently, so the two triangles of a diagonal tile would race.
struct EpiNatP : EpiBase {
    double* P;
    double* work;
    double* G;
    int ld, Tp, T;
    double gamma;
    __device__ void operator()(int b, int gr, int gc, double v0, double v1) {
        size_t off = (size_t)b * Tp * ld + (size_t)gr * ld + gc;
        double v[2] = {v0, v1}, o[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int c = gc + e;
            double eye = gr == c ? 1.0 : 0.0;
            if (gr < T && c < T) {
                o[e] = (1.0 - gamma) * P[off + e] + gamma * (eye + v[e]);
            } else {
                o[e] = eye;
            }
            if (G && c <= gr) {
                const double gval = (gr < T && c < T) ? v[e] : 0.0;
                double* gb = G + (size_t)b * Tp * ld;
                gb[(size_t)gr * ld + c] = gval;
                gb[(size_t)c * ld + gr] = gval;
            }
        }
        *reinterpret_cast<double2*>(P + off) = make_double2(o[0], o[1]);
        *reinterpret_cast<double2*>(work + off) = make_double2(o[0], o[1]);
    }
};

// Phi = tril(L^T Lbar) with the diagonal halved (Cholesky reverse-mode, Murray 2016), Lbar = r q_mu^T - D^-1 L S
// (d ELBO / d L).  L^T Lbar = (L^T r) q_mu^T - (L^T D^-1 L) S = v q_mu^T - G S with the G the natural-gradient half
// formed anyway (EpiNatP): ONE symmetric-by-symmetric product on the lower tiles (T^3 flops) where round 1 formed
// S L^T (T^3) and then L^T Lbar (T^3 / 3).  The accumulators are (G S)[gr, c].
struct EpiPhiGS : EpiBase {
    double* out;
    const double* v;     // L^T r  [B, T]
    const double* q_mu;  // [B, T]
    int ld, Tp, T;
    __device__ void operator()(int b, int gr, int gc, double a0, double a1) {
        double a[2] = {a0, a1}, o[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int c = gc + e;
            if (gr < T && c < T && c <= gr) {
                const double f = v[(size_t)b * T + gr] * q_mu[(size_t)b * T + c] - a[e];
                o[e] = c == gr ? 0.5 * f : f;
            } else {
                o[e] = 0.0;
            }
        }
        *reinterpret_cast<double2*>(out + (size_t)b * Tp * ld + (size_t)gr * ld + gc) = make_double2(o[0], o[1]);
    }
};

// g_theta = sum_{a,b} Kbar_u[a,b] dK[a,b]/dtheta for theta = (variance, lengthscale).  Kbar_u (a plain GEMM result,
// padded [Tp, ld]) is read once; the Matern terms are recomputed from X exactly as k_matern32 computes the gram
// (same staging: the tile's rows and columns of X / l k-major in shared memory with their squared norms, 2 x 2 entries
// at a time, branch-free sqrt_pos / exp_neg).  One CTA per 128 x 128 tile and problem; per-CTA partial sums go to
// partial[b][tile][2] and are added in a fixed order by k_vgp_adam.
// Round 1 did this in the EPILOGUE of the GEMM that forms Kbar_u (so that Kbar_u was never stored): there every
// element divided both points' coordinates by l and called the library sqrt and exp (~600 instructions per element
// at R = 10) in a kernel that keeps 8 warps per SM -- 36 % of an L2 iteration at T = 251 and 12 % at T = 3012 (ncu launch
// lists profiles/r02F_vgp_small_launches.md, r02G_vgp_cfg2_launches.md).  Storing Kbar_u costs 16 T^2 bytes of traffic
// per member and iteration, a fraction of a percent of the iteration.
__global__ void __launch_bounds__(256, BE_MATERN_CTAS)
    k_kbar_grad(const double* __restrict__ Kb, int ld, int Tp, const double* __restrict__ X, int B, int T, int R,
                const double* __restrict__ variance, const double* __restrict__ lengthscale,
                double* __restrict__ partial, int nt) {
    extern __shared__ __align__(16) double sm[];
    double* xi = sm;              // [R][128]  rows of the tile, k-major
    double* xj = xi + NB * R;     // [R][128]  columns of the tile
    double* si = xj + NB * R;     // |x_i / l|^2
    double* sj = si + NB;
    __shared__ double red[8][2];
    const int tile = blockIdx.x / B, b = blockIdx.x % B;
    const int ti = tile / nt, tj = tile % nt;
    const double ls = lengthscale[b];
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
        double q = 0.0;
        for (int k = 0; k < R; ++k) q += src[k * NB + r] * src[k * NB + r];
        (threadIdx.x < NB ? si : sj)[r] = q;
    }
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const double* kb = Kb + (size_t)b * Tp * ld;
    double g0 = 0.0, g1 = 0.0;
#pragma unroll 1
    for (int ip = 0; ip < 4; ++ip) {
        const int lr = ty * 8 + 2 * ip;
        const int gi0 = ti * NB + lr;
        const double2 si2 = *reinterpret_cast<const double2*>(si + lr);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            const int lc = 2 * tx + 32 * c;
            const int gj = tj * NB + lc;
            if (gi0 >= T || gj >= T) continue;
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
            // the padded buffer has zeros beyond T (EpiStore), so the pair loads need no column guard
            const double2 k0 = *reinterpret_cast<const double2*>(kb + (size_t)gi0 * ld + gj);
            const double2 k1 = gi0 + 1 < T ? *reinterpret_cast<const double2*>(kb + (size_t)(gi0 + 1) * ld + gj)
                                           : make_double2(0.0, 0.0);
            const double kv[2][2] = {{k0.x, k0.y}, {k1.x, k1.y}};
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (gi0 + i >= T || gj + e >= T) continue;
                    const double r2 = (-2.0 * acc[i][e] + (i ? si2.y : si2.x)) + (e ? sj2.y : sj2.x);
                    const double x = SQRT3 * sqrt_pos(fmax(r2, 1e-36));
                    const double ex = x <= 700.0 ? exp_neg(x) : exp(-x);
                    g0 = fma(kv[i][e], (1.0 + x) * ex, g0);
                    if (r2 > 1e-36) g1 = fma(kv[i][e], (x * x) * ex, g1);  // x = sqrt3 r: 3 r^2 e^(-sqrt3 r)
                }
        }
    }
    g1 *= variance[b] / ls;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        g0 += __shfl_xor_sync(0xffffffffu, g0, o);
        g1 += __shfl_xor_sync(0xffffffffu, g1, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[warp][0] = g0;
        red[warp][1] = g1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int w = 0; w < 8; ++w) {
            s0 += red[w][0];
            s1 += red[w][1];
        }
        partial[((size_t)b * nt * nt + tile) * 2] = s0;
        partial[((size_t)b * nt * nt + tile) * 2 + 1] = s1;
    }
}

// cov = K + AT (S - I) AT^T + D   (predict_f full_cov + models.py:220), lower tiles, mirrored
struct EpiCov : EpiBase {
    const double* K;  // padded [Tp, ld], both triangles, no jitter
    const double* y_var;
    double* cov;  // dense [B, T, T]
    double* var_diag;
    int ld, Tp, T;
    __device__ void operator()(int b, int gr, int gc, double v0, double v1) {
        if (gr >= T) return;
        double v[2] = {v0, v1};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int c = gc + e;
            if (c >= T || c > gr) continue;
            double val = K[(size_t)b * Tp * ld + (size_t)gr * ld + c] + v[e];
            if (c == gr) {
                val += y_var[(size_t)b * T + gr];
                var_diag[(size_t)b * T + gr] = val;
            }
            cov[(size_t)b * T * T + (size_t)gr * T + c] = val;
            cov[(size_t)b * T * T + (size_t)c * T + gr] = val;
        }
    }
};

// out[a][i] = in[i][a] restricted to one triangle of `in` (tri = 1: lower, a <= i; tri = 2: upper,
// a >= i; 0: everything), optionally scaled by colscale[i] (0 on padding).  32 x 32 smem tiles.
__global__ void __launch_bounds__(256) k_transpose(const double* __restrict__ in, int ld, int Tp, int T, int tri,
                                                   double* __restrict__ out, double* __restrict__ out_scaled,
                                                   const double* __restrict__ y_var, int B) {
    __shared__ double tile[32][33];
    const int nt = (Tp + 31) / 32;
    int t = blockIdx.x % (nt * nt), b = blockIdx.x / (nt * nt);
    const int ti = t / nt, ta = t % nt;  // input rows block ti, input cols block ta
    const double* ib = in + (size_t)b * Tp * ld;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int rr = ty; rr < 32; rr += 8) {
        int i = ti * 32 + rr, a = ta * 32 + tx;
        double v = 0.0;
        if (i < Tp && a < Tp && (tri == 0 || (tri == 1 ? a <= i : a >= i))) v = ib[(size_t)i * ld + a];
        tile[rr][tx] = v;
    }
    __syncthreads();
    for (int rr = ty; rr < 32; rr += 8) {
        int a = ta * 32 + rr, i = ti * 32 + tx;
        if (a < Tp && i < Tp) {
            double v = tile[tx][rr];
            out[(size_t)b * Tp * ld + (size_t)a * ld + i] = v;
            if (out_scaled) out_scaled[(size_t)b * Tp * ld + (size_t)a * ld + i] = i < T ? v / y_var[(size_t)b * T + i] : 0.0;
        }
    }
}

// One warp per row: s = sum_k M[i][k] x[k] over the triangle given by tri (0 full, 1 k <= i, 2 k >= i).
//   op 0: out[i] = s
//   op 1: out[i] = (1-gamma) out[i] + gamma s, with x[k] := y[k] / y_var[k]   (theta_1 step)
//   op 2: out[i] = (y[i] - s) / y_var[i]                                      (residual r)
__global__ void __launch_bounds__(256) k_rowdot(const double* __restrict__ M, int ld, int Tp, int T, int tri, int op,
                                                const double* __restrict__ x, const double* __restrict__ y,
                                                const double* __restrict__ y_var, double gamma,
                                                double* __restrict__ out, int B) {
    int wg = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (wg >= B * T) return;
    int b = wg / T, i = wg % T;
    const double* row = M + (size_t)b * Tp * ld + (size_t)i * ld;
    int k0 = tri == 2 ? i : 0, k1 = tri == 1 ? i + 1 : T;
    double s = 0.0;
    for (int k = k0 + lane; k < k1; k += 32) {
        double xv = op == 1 ? y[(size_t)b * T + k] / y_var[(size_t)b * T + k] : x[(size_t)b * T + k];
        s = fma(row[k], xv, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        size_t gi = (size_t)b * T + i;
        if (op == 0) out[gi] = s;
        else if (op == 1) out[gi] = (1.0 - gamma) * out[gi] + gamma * s;
        else out[gi] = (y[gi] - s) / y_var[gi];
    }
}

__device__ __forceinline__ double softplus_d(double u) { return u > 0.0 ? u + log1p(exp(-u)) : log1p(exp(u)); }

// u [B,2] unconstrained (variance, lengthscale) -> constrained values (gpflow's softplus bijector)
__global__ void k_vgp_constrain(const double* __restrict__ u, int B, double* __restrict__ variance,
                                double* __restrict__ lengthscale) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    variance[b] = softplus_d(u[2 * b]);
    lengthscale[b] = softplus_d(u[2 * b + 1]);
}

// tf.optimizers.Adam (TF 2.8 OptimizerV2) step on the two unconstrained parameters (models.py:192,210):
// g_u = -(dELBO/dtheta) * dsoftplus/du,  dsoftplus/du = 1 - exp(-theta);  then the new constrained values.
__global__ void k_vgp_adam(const double* __restrict__ partial, int ctas_per_problem, int B, double lr, double b1,
                           double b2, double eps, double* __restrict__ u, double* __restrict__ m,
                           double* __restrict__ v, int* __restrict__ step, double* __restrict__ variance,
                           double* __restrict__ lengthscale) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double g[2] = {0.0, 0.0};
    for (int c = 0; c < ctas_per_problem; ++c) {
        g[0] += partial[((size_t)b * ctas_per_problem + c) * 2];
        g[1] += partial[((size_t)b * ctas_per_problem + c) * 2 + 1];
    }
    const double theta[2] = {variance[b], lengthscale[b]};
    const int t = step[b] + 1;
    step[b] = t;
    const double lr_t = lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t));
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        double gu = -g[p] * (-expm1(-theta[p]));
        double mm = b1 * m[2 * b + p] + (1.0 - b1) * gu;
        double vv = b2 * v[2 * b + p] + (1.0 - b2) * gu * gu;
        m[2 * b + p] = mm;
        v[2 * b + p] = vv;
        u[2 * b + p] -= lr_t * mm / (sqrt(vv) + eps);
    }
    variance[b] = softplus_d(u[2 * b]);
    lengthscale[b] = softplus_d(u[2 * b + 1]);
}

// constrained (variance, lengthscale) -> unconstrained u = softplus^-1(theta) = theta + log(1 - exp(-theta))
__global__ void k_vgp_unconstrain(const double* __restrict__ variance, const double* __restrict__ lengthscale, int B,
                                  double* __restrict__ u) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    u[2 * b] = variance[b] + log(-expm1(-variance[b]));
    u[2 * b + 1] = lengthscale[b] + log(-expm1(-lengthscale[b]));
}

// padded identity (q_sqrt = I  =>  P = S = I)
__global__ void k_set_identity(double* __restrict__ M, int ld, int Tp, int B) {
    size_t n = (size_t)B * Tp * ld;
    for (size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gid < n; gid += (size_t)gridDim.x * blockDim.x) {
        size_t rem = gid % ((size_t)Tp * ld);
        M[gid] = (rem / ld == rem % ld) ? 1.0 : 0.0;
    }
}

}  // namespace be
