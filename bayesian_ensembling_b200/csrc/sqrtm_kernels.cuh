// Kernels of the matrix-square-root family: sqrtm (ensembles/wasserstein.py:10-13), the Gaussian
// W2 "distance" (wasserstein.py:21-47) and the full-covariance barycentre fixed point
// (BASELINE config 5; definition in DESIGN.md 3.4).
//
// The reference's sqrtm is U diag(sqrt(s)) V^H from an SVD; for the symmetric positive
// definite matrices the path feeds it that IS the principal square root.  Here it is computed
// with the scaled Denman-Beavers iteration
//     Y <- (mu Y + Z^-1 / mu) / 2,   Z <- (mu Z + Y^-1 / mu) / 2,   Y0 = A, Z0 = I,
//     mu = |det Y det Z|^(-1/2T)  (determinant scaling, switched off once the step is small)
// whose iterates stay SPD, so every step is two SPD inversions = potrf + trtri + lauum on the
// FP64 tensor-core tile engine (be_api.cu), and Y -> A^1/2, Z -> A^-1/2 quadratically
// (7 iterations to ~1e-15 on posterior covariances of condition number 1e2..1e3).
#pragma once
#include "vgp_kernels.cuh"

namespace be {

// dense symmetric [B,T,T] -> padded [B,Tp,ld], BOTH triangles, identity padding
__global__ void k_pad_full(const double* __restrict__ A, int B, int T, int Tp, int ld, double* __restrict__ W,
                           double* __restrict__ W2) {
    size_t n = (size_t)B * Tp * ld;
    for (size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gid < n; gid += (size_t)gridDim.x * blockDim.x) {
        int b = (int)(gid / ((size_t)Tp * ld));
        size_t rem = gid % ((size_t)Tp * ld);
        int i = (int)(rem / ld), j = (int)(rem % ld);
        double v = (i < T && j < T) ? A[(size_t)b * T * T + (size_t)i * T + j] : (i == j ? 1.0 : 0.0);
        W[gid] = v;
        if (W2) W2[gid] = v;
    }
}

// padded (both triangles) -> dense [B,T,T]
__global__ void k_copy_out_full(const double* __restrict__ Work, int ld, int Tp, int T, double* __restrict__ out, int B) {
    size_t n = (size_t)B * T * T;
    for (size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gid < n; gid += (size_t)gridDim.x * blockDim.x) {
        int b = (int)(gid / ((size_t)T * T));
        size_t rem = gid % ((size_t)T * T);
        int i = (int)(rem / T), j = (int)(rem % T);
        out[gid] = Work[(size_t)b * Tp * ld + (size_t)i * ld + j];
    }
}

// out[b] = scale * I (padded; the padding diagonal is 1)
__global__ void k_set_scaled_identity(double* __restrict__ M, int ld, int Tp, int T, int B, double scale) {
    size_t n = (size_t)B * Tp * ld;
    for (size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gid < n; gid += (size_t)gridDim.x * blockDim.x) {
        size_t rem = gid % ((size_t)Tp * ld);
        int i = (int)(rem / ld), j = (int)(rem % ld);
        M[gid] = i == j ? (i < T ? scale : 1.0) : 0.0;
    }
}

// out[b] (+)= sum_{j<T} f(W[b][j][j]) with f = log (of |.|, Cholesky diagonals: half log-det) or identity (trace).
// One CTA per problem, fixed-order reduction.
template <int LOG>
__global__ void __launch_bounds__(256) k_diag_reduce(const double* __restrict__ W, int ld, int Tp, int T,
                                                     double* __restrict__ out, int accumulate) {
    __shared__ double red[8];
    const double* Wb = W + (size_t)blockIdx.x * Tp * ld;
    double s = 0.0;
    for (int j = threadIdx.x; j < T; j += 256) {
        double d = Wb[(size_t)j * ld + j];
        s += LOG ? log(fabs(d)) : d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        out[blockIdx.x] = accumulate ? out[blockIdx.x] + t : t;
    }
}

// mu[b] = exp(-(2 hld_y + 2 hld_z) / (2T)) while scaling is on for problem b (delta[b] >= switch_off), else 1
__global__ void k_db_mu(const double* __restrict__ hld_y, const double* __restrict__ hld_z, const double* __restrict__ delta,
                        double switch_off, int T, int B, double* __restrict__ mu) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    mu[b] = delta[b] >= switch_off ? exp(-(hld_y[b] + hld_z[b]) / (double)T) : 1.0;
}

// Denman-Beavers combine on the lower tiles of inv = V V^T (the accumulators):
//   cur[i,j] <- (mu cur[i,j] + inv[i,j] / mu) / 2        (in place: only j <= i is read)
// written to both triangles of `cur` and, lower triangle + identity padding, to `work` (the
// buffer the next factorisation destroys).  Accumulates |new - old|^2 and |new|^2 per CTA.
struct EpiDB {
    double* cur;
    double* work;
    const double* mu;
    double* partial;  // [B][ctas][2]
    int ld, Tp, T, ctas_per_problem;
    double d2, n2;
    __device__ void operator()(int b, int gr, int gc, double v0, double v1) {
        const double m = mu[b];
        double* cb = cur + (size_t)b * Tp * ld;
        double* wb = work + (size_t)b * Tp * ld;
        double v[2] = {v0, v1};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int c = gc + e;
            if (c > gr) continue;
            double val;
            if (gr < T) {  // c <= gr < T
                double old = cb[(size_t)gr * ld + c];
                val = 0.5 * (m * old + v[e] / m);
                double d = val - old;
                double wgt = c == gr ? 1.0 : 2.0;
                d2 += wgt * d * d;
                n2 += wgt * val * val;
            } else {
                val = gr == c ? 1.0 : 0.0;
            }
            cb[(size_t)gr * ld + c] = val;
            cb[(size_t)c * ld + gr] = val;
            wb[(size_t)gr * ld + c] = val;
        }
    }
    __device__ void finish(int b, int cta, double* red) {
        __syncthreads();
        double a0 = d2, a1 = n2;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) {
            red[2 * warp] = a0;
            red[2 * warp + 1] = a1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s0 = 0, s1 = 0;
            for (int w = 0; w < GEMM_THREADS / 32; ++w) {
                s0 += red[2 * w];
                s1 += red[2 * w + 1];
            }
            partial[((size_t)b * ctas_per_problem + cta) * 2] = s0;
            partial[((size_t)b * ctas_per_problem + cta) * 2 + 1] = s1;
        }
    }
};

// delta[b] = |Y_new - Y_old|_F / |Y_new|_F from the per-CTA partial sums (fixed order)
__global__ void k_db_delta(const double* __restrict__ partial, int ctas_per_problem, int B, double* __restrict__ delta) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double d2 = 0.0, n2 = 0.0;
    for (int c = 0; c < ctas_per_problem; ++c) {
        d2 += partial[((size_t)b * ctas_per_problem + c) * 2];
        n2 += partial[((size_t)b * ctas_per_problem + c) * 2 + 1];
    }
    delta[b] = n2 > 0.0 ? sqrt(d2 / n2) : (d2 > 0.0 ? INFINITY : 0.0);
}

// a problem that leaves the Denman-Beavers loop above the tolerance (max_iters reached, or a NaN iterate that the
// Cholesky did not already report) is marked BE_INFO_NOT_CONVERGED; Cholesky reports (> 0) are kept
__global__ void k_mark_unconverged(const double* __restrict__ delta, double tol, int B, int* __restrict__ info) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (!(delta[b] < tol) && info[b] == 0) info[b] = BE_INFO_NOT_CONVERGED;
}

// symmetric product computed on lower tiles: out[i,j] = out[j,i] = acc (j <= i), identity padding
struct EpiSym : EpiBase {
    double* out;
    double* work;  // optional second copy (lower triangle is what a factorisation reads)
    int ld, Tp, T;
    __device__ void operator()(int b, int gr, int gc, double v0, double v1) {
        double* ob = out + (size_t)b * Tp * ld;
        double v[2] = {v0, v1};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int c = gc + e;
            if (c > gr) continue;
            double val = gr < T ? v[e] : (gr == c ? 1.0 : 0.0);
            ob[(size_t)gr * ld + c] = val;
            ob[(size_t)c * ld + gr] = val;
            if (work) work[(size_t)b * Tp * ld + (size_t)gr * ld + c] = val;
        }
    }
};

// plain product: out[i,j] = acc on the real part, 0 on the padding (operand of a further product)
struct EpiPlain : EpiBase {
    double* out;
    int ld, Tp, T;
    __device__ void operator()(int b, int gr, int gc, double v0, double v1) {
        double o0 = (gr < T && gc < T) ? v0 : 0.0;
        double o1 = (gr < T && gc + 1 < T) ? v1 : 0.0;
        *reinterpret_cast<double2*>(out + (size_t)b * Tp * ld + (size_t)gr * ld + gc) = make_double2(o0, o1);
    }
};

// cand[c] = sum_m w[c,m] Q[c,m]  (sequential over m, the order of wasserstein.py:85-86), for the
// cells whose fixed point is still running; both padded copies (S and the factorisation work
// buffer) are refreshed.  Grid-stride over C * Tp * ld.
__global__ void k_weighted_sum(const double* __restrict__ Q, const double* __restrict__ w, const int* __restrict__ active,
                               int C, int M, int Tp, int ld, int T, double* __restrict__ S, double* __restrict__ S2) {
    const size_t per = (size_t)Tp * ld;
    size_t n = (size_t)C * per;
    for (size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gid < n; gid += (size_t)gridDim.x * blockDim.x) {
        int c = (int)(gid / per);
        if (!active[c]) continue;
        size_t rem = gid % per;
        int i = (int)(rem / ld), j = (int)(rem % ld);
        double v;
        if (i < T && j < T) {
            v = 0.0;
            for (int m = 0; m < M; ++m) v += w[(size_t)c * M + m] * Q[((size_t)c * M + m) * per + rem];
        } else {
            v = i == j ? 1.0 : 0.0;
        }
        S[gid] = v;
        if (S2) S2[gid] = v;
    }
}

// mu[c, t] = sum_m w[c,m] mus[c,m,t]   (wasserstein.py:98)
__global__ void k_weighted_mean(const double* __restrict__ mus, const double* __restrict__ w, int C, int M, int T,
                                double* __restrict__ out) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)C * T) return;
    int c = (int)(gid / T), t = (int)(gid % T);
    double s = 0.0;
    for (int m = 0; m < M; ++m) s += w[(size_t)c * M + m] * mus[((size_t)c * M + m) * T + t];
    out[gid] = s;
}

// w2[p] = |mu1 - mu2|_2 + tr(S1) + tr(S2) - 2 tr(Q)   (wasserstein.py:40-45; the location term is
// NOT squared, quirk Q-W2).  One CTA per pair; the traces come in as sums already.
__global__ void __launch_bounds__(256) k_w2_finish(const double* __restrict__ mu1, const double* __restrict__ mu2, int T,
                                                   const double* __restrict__ tr1, const double* __restrict__ tr2,
                                                   const double* __restrict__ trq, double* __restrict__ w2) {
    __shared__ double red[8];
    int p = blockIdx.x;
    double s = 0.0;
    for (int j = threadIdx.x; j < T; j += 256) {
        double d = mu1[(size_t)p * T + j] - mu2[(size_t)p * T + j];
        s += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        w2[p] = sqrt(t) + ((tr1[p] + tr2[p]) - 2.0 * trq[p]);
    }
}

// full_cov=False branch of wasserstein.py:36-45: the covariances are diag(variance), for which
// sqrtm is the elementwise square root:  w2 = |mu1-mu2| + sum_j (v1 + v2 - 2 sqrt(sqrt(v1) v2 sqrt(v1)))
__global__ void __launch_bounds__(256) k_w2_diag(const double* __restrict__ mu1, const double* __restrict__ var1,
                                                 const double* __restrict__ mu2, const double* __restrict__ var2, int T,
                                                 double* __restrict__ w2) {
    __shared__ double red[2][8];
    int p = blockIdx.x;
    double s = 0.0, g = 0.0;
    for (int j = threadIdx.x; j < T; j += 256) {
        size_t o = (size_t)p * T + j;
        double d = mu1[o] - mu2[o];
        s += d * d;
        double r1 = sqrt(var1[o]);
        g += (var1[o] + var2[o]) - 2.0 * sqrt(r1 * var2[o] * r1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        g += __shfl_xor_sync(0xffffffffu, g, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s;
        red[1][threadIdx.x >> 5] = g;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0, u = 0.0;
        for (int w = 0; w < 8; ++w) {
            t += red[0][w];
            u += red[1][w];
        }
        w2[p] = sqrt(t) + u;
    }
}

}  // namespace be
