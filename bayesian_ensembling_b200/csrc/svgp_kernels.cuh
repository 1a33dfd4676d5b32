// Kernels of the SVGP stage of GPDTW3D.fit (ensembles/models.py:357-411): a sparse variational GP over
// all (t, lat, lon) points with a sum of four Matern-3/2 kernels, M inducing inputs (400 in the reference), trained
// by alternating a natural-gradient step on (q_mu, q_sqrt) and an Adam step on the kernel parameters AND the
// inducing inputs, each on its own minibatch (500 points).  Every matrix here is at most M x M or M x batch
// (400 x 500): the M x M factorisations run on the blocked DMMA path of be_kernels.cuh (one problem), the
// rectangular products on a warp-per-tile DMMA GEMM that reads its fragments straight from global memory (k_dgemm) -- at
// these sizes a step is bound by the latency of the single-problem factorisation kernels, not by the FP64 pipe, and the
// stage is a "next" row, not the headline.
// The arithmetic follows oracle/svgp.py line by line (GPflow 2.1.5's SVGP with whiten=True, num_data=None).
#pragma once
#include <math.h>

#include "be_kernels.cuh"

namespace be {

constexpr int SVGP_COMPONENTS = 4;
constexpr int SVGP_MAX_D = 36;  // 4 + realisations

struct SvgpKernelParams {
    const double* variance;     // [4] device
    const double* lengthscale;  // [4] device
    int D;                      // 4 + R: columns (x, y, z, t, realisations...)  (models.py:270-319)
};

// active dimensions of component c in the order of models.py:358-364: time [3], (x, y) [0, 1], z [2], realisations [4, D)
__device__ __forceinline__ void svgp_dims(int c, int D, int& d0, int& d1) {
    if (c == 0) { d0 = 3; d1 = 4; }
    else if (c == 1) { d0 = 0; d1 = 2; }
    else if (c == 2) { d0 = 2; d1 = 3; }
    else { d0 = 4; d1 = D; }
}

// r2 of component c between rows a and b in GPflow's expansion form (|a/l|^2 + |b/l|^2 - 2 a.b / l^2)
__device__ __forceinline__ double svgp_r2(const double* __restrict__ a, const double* __restrict__ b, int d0, int d1, double ls) {
    double dot = 0.0, sa = 0.0, sb = 0.0;
    for (int d = d0; d < d1; ++d) {
        const double x = a[d] / ls, y = b[d] / ls;
        dot = fma(x, y, dot);
        sa = fma(x, x, sa);
        sb = fma(y, y, sb);
    }
    return (-2.0 * dot + sa) + sb;
}

// K[i, j] = sum_c s2_c (1 + sqrt3 r_c) exp(-sqrt3 r_c) between rows i of A [na, D] and j of Bm [nb, D]; out row-major
// with leading dimension ld.  diag_add is added where i == j (the jitter of Kuu); if pad_to > na the rows / columns
// [na, pad_to) are identity padding (the padded layout of be_kernels.cuh: out must then be [pad_to, ld]).
__global__ void k_svgp_kernel(const double* __restrict__ A, int na, const double* __restrict__ Bm, int nb, SvgpKernelParams p,
                              double diag_add, int pad_to, double* __restrict__ out, int ld) {
    const int rows = pad_to > na ? pad_to : na, cols = pad_to > na ? pad_to : nb;
    const size_t n = (size_t)rows * cols;
    for (size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gid < n; gid += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(gid / cols), j = (int)(gid % cols);
        double v;
        if (i < na && j < nb) {
            v = (i == j) ? diag_add : 0.0;
            for (int c = 0; c < SVGP_COMPONENTS; ++c) {
                int d0, d1;
                svgp_dims(c, p.D, d0, d1);
                const double r = sqrt(fmax(svgp_r2(A + (size_t)i * p.D, Bm + (size_t)j * p.D, d0, d1, p.lengthscale[c]), 1e-36));
                v += p.variance[c] * (1.0 + SQRT3 * r) * exp(-SQRT3 * r);
            }
        } else {
            v = i == j ? 1.0 : 0.0;
        }
        out[(size_t)i * ld + j] = v;
    }
}

// FP64 GEMM for the rectangular products of the step (every dimension <= ~500, any of the three layouts):
//   C [m, n] = alpha op(A) op(B) + beta C, row-major;  op(A) [m, k]: A[i * lda + kk] (TA = 0) or A[kk * lda + i] (TA = 1),
//   op(B) [k, n]: B[kk * ldb + j] (TB = 0) or B[j * ldb + kk] (TB = 1).
// One WARP per 16 x 32 output tile, DMMA m8n8k4 fragments read straight from global memory (the operands are L2 /
// L1-resident at these sizes; each 8-byte lane load is one fragment element, so any layout costs the same), four
// k-steps per register group and the next group's 24 loads in flight under the current group's 32 DMMAs.  No shared
// memory, no barrier: the first version (64 x 64 shared-memory tiles, 4 x 4 outputs per thread) put 56 CTAs on 148 SMs
// and took ~140 us per product (1.1 TFLOP/s, 48 % of the step); 400 independent warps are spread over all of them.
// Sums run over k in order within a lane's accumulator, as DMMA does: deterministic.
template <int TA, int TB>
__global__ void __launch_bounds__(128) k_dgemm(int m, int n, int k, double alpha, const double* __restrict__ A, int lda,
                                               const double* __restrict__ Bm, int ldb, double beta, double* __restrict__ C, int ldc) {
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int tiles_n = (n + 31) / 32, tiles_m = (m + 15) / 16;
    const int wt = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (wt >= tiles_m * tiles_n) return;
    const int i0 = (wt / tiles_n) * 16, j0 = (wt % tiles_n) * 32;
    constexpr int U = 4;  // k-steps (of 4) per register group
    double acc[2][4][2] = {};
    double a0[U][2], b0[U][4], a1[U][2], b1[U][4];
    auto load_group = [&](double (&a)[U][2], double (&b)[U][4], int k0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int kk = k0 + 4 * u + q;
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const int i = i0 + 8 * mi + g;
                a[u][mi] = (i < m && kk < k) ? (TA ? A[(size_t)kk * lda + i] : A[(size_t)i * lda + kk]) : 0.0;
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int j = j0 + 8 * ni + g;
                b[u][ni] = (j < n && kk < k) ? (TB ? Bm[(size_t)j * ldb + kk] : Bm[(size_t)kk * ldb + j]) : 0.0;
            }
        }
    };
    auto mma_group = [&](const double (&a)[U][2], const double (&b)[U][4]) {
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[u][mi], b[u][ni]);
    };
    load_group(a0, b0, 0);
    for (int k0 = 0; k0 < k; k0 += 8 * U) {
        load_group(a1, b1, k0 + 4 * U);  // out-of-range k loads nothing and contributes zeros
        mma_group(a0, b0);
        load_group(a0, b0, k0 + 8 * U);
        mma_group(a1, b1);
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gi = i0 + 8 * mi + g, gj = j0 + 8 * ni + 2 * q + e;
                if (gi < m && gj < n) {
                    double* c = C + (size_t)gi * ldc + gj;
                    *c = beta == 0.0 ? alpha * acc[mi][ni][e] : fma(alpha, acc[mi][ni][e], beta * *c);
                }
            }
}
inline unsigned dgemm_grid(int m, int n) { return (unsigned)((((m + 15) / 16) * ((n + 31) / 32) + 3) / 4); }

// y [m] = A [m, k] x  (warp per row)
__global__ void k_gemv_n(int m, int k, const double* __restrict__ A, int lda, const double* __restrict__ x, double* __restrict__ y) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= m) return;
    double s = 0.0;
    for (int j = lane; j < k; j += 32) s = fma(A[(size_t)row * lda + j], x[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] = s;
}
// y [n] = A^T x, A [m, n]: 32 columns per CTA, the rows dealt over 8 row groups (a thread per column walking all m rows
// was a chain of m dependent loads: 89 us at m = 400), partial sums combined in a fixed order.  Launch with 256 threads.
__global__ void __launch_bounds__(256) k_gemv_t(int m, int n, const double* __restrict__ A, int lda,
                                                const double* __restrict__ x, double* __restrict__ y) {
    __shared__ double part[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    double s = 0.0;
    if (j < n)
        for (int i = ty; i < m; i += 8) s = fma(A[(size_t)i * lda + j], x[i], s);
    part[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && j < n) {
        double t = part[0][tx];
#pragma unroll
        for (int r = 1; r < 8; ++r) t += part[r][tx];
        y[j] = t;
    }
}

// minibatch gather: Xb [n, D], yb [n], sb [n] from X [N, D], Y [N, 2] (columns: DTW mean, variance; models.py:180)
// The minibatch is row (2 * *step + half) of idx [2 * n_steps, n]: the step counter lives on the device so that one
// optimisation step can be captured in a CUDA graph and replayed.
__global__ void k_svgp_gather(const double* __restrict__ X, const double* __restrict__ Y, const long long* __restrict__ idx,
                              const int* __restrict__ step, int half, int n, int D, double* __restrict__ Xb,
                              double* __restrict__ yb, double* __restrict__ sb) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * D) return;
    const int i = gid / D, d = gid % D;
    const long long src = idx[(size_t)(2 * *step + half) * n + i];
    Xb[gid] = X[(size_t)src * D + d];
    if (d == 0) {
        yb[i] = Y[(size_t)src * 2];
        sb[i] = Y[(size_t)src * 2 + 1];
    }
}

// natural-gradient targets: Aw = A / s (columnwise) for nat1* = Aw y and P* = I + Aw A^T
__global__ void k_svgp_scale_cols(const double* __restrict__ A, const double* __restrict__ s, int M, int n, double* __restrict__ Aw) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)M * n) return;
    Aw[gid] = A[gid] / s[gid % n];
}
// P <- (1 - gamma) P + gamma (I + G), n1 <- (1 - gamma) n1 + gamma n1*; P in the padded [Mp, Mp] layout, G [M, M] (ld M);
// Work receives a copy of the new P (symmetrised as the oracle does) with identity padding, to be factored in place.
__global__ void k_svgp_natgrad_update(double* __restrict__ P, const double* __restrict__ G, double* __restrict__ n1,
                                      const double* __restrict__ n1s, int M, int Mp, double gamma, double* __restrict__ Work) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)Mp * Mp) return;
    const int i = (int)(gid / Mp), j = (int)(gid % Mp);
    double v;
    if (i < M && j < M) {
        const double gij = 0.5 * (G[(size_t)i * M + j] + G[(size_t)j * M + i]);
        v = (1.0 - gamma) * P[gid] + gamma * ((i == j ? 1.0 : 0.0) + gij);
        P[gid] = v;
    } else {
        v = i == j ? 1.0 : 0.0;
    }
    Work[gid] = v;
    if (j == 0 && i < M) n1[i] = (1.0 - gamma) * n1[i] + gamma * n1s[i];
}

// copy the top-left M x M of a padded matrix into a padded work buffer with identity padding (symmetric source)
__global__ void k_svgp_pad_copy(const double* __restrict__ S, int lds, int M, int Mp, double* __restrict__ Work) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)Mp * Mp) return;
    const int i = (int)(gid / Mp), j = (int)(gid % Mp);
    Work[gid] = (i < M && j < M) ? S[(size_t)i * lds + j] : (i == j ? 1.0 : 0.0);
}

// per-point likelihood terms at the minibatch: gm = (y - m) / s, gv = -1 / (2 s)   (oracle/svgp.py:elbo_grads)
__global__ void k_svgp_point_grads(const double* __restrict__ fmean, const double* __restrict__ y, const double* __restrict__ s,
                                   int n, double* __restrict__ gm, double* __restrict__ gv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    gm[i] = (y[i] - fmean[i]) / s[i];
    gv[i] = -0.5 / s[i];
}
// Abar = q_mu gm^T + 2 (Sq Sq^T A - A) diag(gv)
__global__ void k_svgp_abar(const double* __restrict__ A, const double* __restrict__ SW, const double* __restrict__ q_mu,
                            const double* __restrict__ gm, const double* __restrict__ gv, int M, int n, double* __restrict__ Abar) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)M * n) return;
    const int m = (int)(gid / n), i = (int)(gid % n);
    Abar[gid] = q_mu[m] * gm[i] + 2.0 * (SW[gid] - A[gid]) * gv[i];
}
// Lbar <- -tril(Lbar) in place (M x M, ld M); strict upper part zero
__global__ void k_svgp_neg_tril(double* __restrict__ Lb, int M) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)M * M) return;
    const int i = (int)(gid / M), j = (int)(gid % M);
    Lb[gid] = j <= i ? -Lb[gid] : 0.0;
}
// Phi <- tril(Phi) with the diagonal halved (Cholesky back-propagation, Murray 2016 eq. 10)
__global__ void k_svgp_phi(double* __restrict__ Phi, int M) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)M * M) return;
    const int i = (int)(gid / M), j = (int)(gid % M);
    Phi[gid] = j < i ? Phi[gid] : (j == i ? 0.5 * Phi[gid] : 0.0);
}

// Gradient of the ELBO with respect to the kernel parameters and the inducing inputs from Kuf_bar [M, n] and the
// UNsymmetrised Kuu_bar' [M, M] (the kernel symmetrises: Kuu_bar = (K' + K'^T) / 2).  One CTA per inducing point m:
//   g_var[c] += sum Kbar (1 + sqrt3 r) e,  g_ls[c] += sum Kbar 3 s2 r^2 e / l,
//   g_Z[m, d] += sym * sum_j (-3 s2 e / l^2) Kbar[m, j] (z_md - x_jd)     (sym = 2 for the Kuu term)
// with the r <= 1e-18 clamp passing no gradient; g_var[c] also takes sum_i gv_i (kff = sum_c s2_c) from CTA 0.
// g [8] (variances then lengthscales) is accumulated with atomics and must be zeroed by the caller; g_Z is written.
__global__ void __launch_bounds__(128) k_svgp_param_grads(const double* __restrict__ Z, const double* __restrict__ Xb,
                                                          const double* __restrict__ Kuf_bar, const double* __restrict__ Kuu_bar,
                                                          const double* __restrict__ gv, SvgpKernelParams p, int M, int n,
                                                          double* __restrict__ g, double* __restrict__ g_Z) {
    __shared__ double red[4][SVGP_MAX_D + 8];
    const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.D;
    const double* zm = Z + (size_t)m * D;
    double gvar[SVGP_COMPONENTS] = {}, gls[SVGP_COMPONENTS] = {};
    double gz[SVGP_MAX_D];
    for (int d = 0; d < D; ++d) gz[d] = 0.0;
    for (int j = tid; j < n + M; j += 128) {
        const bool uu = j >= n;
        const int jj = uu ? j - n : j;
        const double* xo = uu ? Z + (size_t)jj * D : Xb + (size_t)jj * D;
        const double kbar = uu ? 0.5 * (Kuu_bar[(size_t)m * M + jj] + Kuu_bar[(size_t)jj * M + m]) : Kuf_bar[(size_t)m * n + jj];
        const double sym = uu ? 2.0 : 1.0;
        for (int c = 0; c < SVGP_COMPONENTS; ++c) {
            int d0, d1;
            svgp_dims(c, D, d0, d1);
            const double ls = p.lengthscale[c], var = p.variance[c];
            const double r2 = svgp_r2(zm, xo, d0, d1, ls);
            const bool live = r2 > 1e-36;
            const double r = sqrt(fmax(r2, 1e-36));
            const double e = exp(-SQRT3 * r);
            gvar[c] += kbar * (1.0 + SQRT3 * r) * e;
            if (live) {
                gls[c] += kbar * 3.0 * var * r * r * e / ls;
                const double coef = sym * kbar * (-3.0 * var * e / (ls * ls));
                for (int d = d0; d < d1; ++d) gz[d] += coef * (zm[d] - xo[d]);
            }
        }
    }
    if (m == 0)
        for (int i = tid; i < n; i += 128)
            for (int c = 0; c < SVGP_COMPONENTS; ++c) gvar[c] += gv[i];
    // block reduction of 8 + D partial sums
    for (int q = 0; q < 8 + D; ++q) {
        double v = q < 4 ? gvar[q] : (q < 8 ? gls[q - 4] : gz[q - 8]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][q] = v;
    }
    __syncthreads();
    for (int q = tid; q < 8 + D; q += 128) {
        const double v = (red[0][q] + red[1][q]) + (red[2][q] + red[3][q]);
        if (q < 8)
            atomicAdd(g + q, v);
        else
            g_Z[(size_t)m * D + (q - 8)] = v;
    }
}

// TF-Keras Adam (beta1 .9, beta2 .999, eps 1e-7, bias-corrected lr) on the loss = -ELBO: the eight unconstrained
// (softplus) kernel parameters u [8] and the inducing inputs Z [M * D]; state am / av sized 8 + M * D; step [1].
__global__ void k_svgp_adam(const double* __restrict__ g, const double* __restrict__ g_Z, int nz, double lr, double* __restrict__ u,
                            double* __restrict__ Z, double* __restrict__ am, double* __restrict__ av, const int* __restrict__ step,
                            double* __restrict__ variance, double* __restrict__ lengthscale) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= 8 + nz) return;
    const double b1 = 0.9, b2 = 0.999, eps = 1e-7;
    const int t = *step + 1;
    double grad;
    if (gid < 8) {
        const double x = gid < 4 ? variance[gid] : lengthscale[gid - 4];
        grad = -g[gid] * (-expm1(-x));  // d softplus(u) / du = 1 - exp(-x)
    } else {
        grad = -g_Z[gid - 8];
    }
    const double mm = b1 * am[gid] + (1.0 - b1) * grad;
    const double vv = b2 * av[gid] + (1.0 - b2) * grad * grad;
    am[gid] = mm;
    av[gid] = vv;
    const double lr_t = lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t));
    const double upd = lr_t * mm / (sqrt(vv) + eps);
    if (gid < 8) {
        const double un = u[gid] - upd;
        u[gid] = un;
        const double sp = un > 0.0 ? un + log1p(exp(-un)) : log1p(exp(un));  // softplus
        if (gid < 4) variance[gid] = sp; else lengthscale[gid - 4] = sp;
    } else {
        Z[gid - 8] -= upd;
    }
}
__global__ void k_svgp_step_inc(int* step) { *step += 1; }
__global__ void k_svgp_unconstrain(const double* __restrict__ variance, const double* __restrict__ lengthscale, double* __restrict__ u) {
    const int i = threadIdx.x;
    if (i >= 8) return;
    const double x = i < 4 ? variance[i] : lengthscale[i - 4];
    u[i] = x + log(-expm1(-x));
}

// predict_f(full_cov=False) at a chunk: fvar_i = kff - sum_m A[m,i]^2 + sum_m W[m,i]^2, W = Sq^T A; out var = fvar + noise
__global__ void k_svgp_predict_var(const double* __restrict__ A, const double* __restrict__ W, const double* __restrict__ variance,
                                   const double* __restrict__ noise, int M, int n, double* __restrict__ var_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = (variance[0] + variance[1]) + (variance[2] + variance[3]);
    double a2 = 0.0, w2 = 0.0;
    for (int m = 0; m < M; ++m) {
        const double a = A[(size_t)m * n + i], w = W[(size_t)m * n + i];
        a2 = fma(a, a, a2);
        w2 = fma(w, w, w2);
    }
    var_out[i] = (s - a2 + w2) + noise[i];
}
__global__ void k_svgp_gather_rows(const double* __restrict__ X, const double* __restrict__ Y, long long start, int n, int D,
                                   double* __restrict__ Xb, double* __restrict__ sb) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * D) return;
    const int i = gid / D, d = gid % D;
    Xb[gid] = X[(size_t)(start + i) * D + d];
    if (d == 0) sb[i] = Y[(size_t)(start + i) * 2 + 1];
}

}  // namespace be
