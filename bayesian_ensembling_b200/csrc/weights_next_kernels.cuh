// SURVEY 8f "next" rows: CRPSWeight (ensembles/weights.py:444-515) and ModelSimilarityWeight
// (weights.py:214-333).  Per-point kernels over [C, M, N] arrays, HBM-bound.
#pragma once
#include "be_kernels.cuh"

namespace be {

constexpr double INV_SQRT_PI = 0.5641895835477563;
constexpr double INV_SQRT_2PI = 0.3989422804014327;

// properscoring.crps_gaussian: sig (z (2 Phi(z) - 1) + 2 phi(z) - 1/sqrt(pi)),  z = (x - mu) / sig.
// 2 Phi(z) - 1 is erf(z / sqrt 2) (no cancellation at small z, and half the cost of normcdf, which is an erfc
// with its own exponential and division); the exponential is exp_tab16 (be_kernels.cuh); the division by sig is
// a product with 1 / sig, formed once per model, when sig is an ordinary number (2^-500 < |sig| < 2^500;
// zero, subnormal, infinite and NaN scales keep the division and with it the reference's inf / NaN results).
__device__ __forceinline__ double crps_gaussian_d(double x, double mu, double sig, double inv_sig, bool ordinary,
                                                  const double* __restrict__ tab) {
    const double d = x - mu;
    const double z = ordinary ? d * inv_sig : d / sig;
    const double a = -0.5 * z * z;
    double e = exp_tab16_core(a, tab);
    if (!exp_tab16_ok(a)) e = exp(a);
    const double pdf = e * INV_SQRT_2PI;
    const double two_cdf_m1 = erf(z * 0.70710678118654752440);
    return sig * (z * two_cdf_m1 + 2.0 * pdf - INV_SQRT_PI);
}

// one thread per (cell, point): weights.py:469-471 (mean over obs realisations), :507 (inverse),
// :510-511 (normalise over models).  scale = the distribution's stddev() -- for the dx.Normal the
// reference builds at :497 that is the member's VARIANCE (quirk Q-SCALE); the caller passes it.
__global__ void k_crps_weights(const double* __restrict__ loc, const double* __restrict__ scale,
                               const double* __restrict__ obs, int C, int M, int Ro, int N, double* __restrict__ w,
                               double* __restrict__ crps_mean, int smem_ok) {
    extern __shared__ double wstage[];
    __shared__ double tab[16];
    if (threadIdx.x < 16) tab[threadIdx.x] = EXP2_16TH[threadIdx.x];
    __syncthreads();
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)C * N) return;
    int c = (int)(gid / N), i = (int)(gid % N);
    WeightStage st(wstage, smem_ok, w + (size_t)c * M * N + i, (size_t)N);
    const double* ob = obs + (size_t)c * Ro * N + i;
    double total = 0.0;
    for (int m = 0; m < M; ++m) {
        size_t o = ((size_t)c * M + m) * N + i;
        const double l = loc[o], sc = scale[o];
        const bool ordinary = fabs(sc) > 0x1p-500 && fabs(sc) < 0x1p500;
        const double inv_sc = 1.0 / sc;
        double s = 0.0;
#pragma unroll 2
        for (int r = 0; r < Ro; ++r) s += crps_gaussian_d(ob[(size_t)r * N], l, sc, inv_sc, ordinary, tab);
        double mean = s / Ro;
        if (crps_mean) crps_mean[o] = mean;
        double inv = 1.0 / mean;
        st[m] = inv;
        if (inv == inv) total += inv;  // xarray .sum('model') skips NaN
    }
    for (int m = 0; m < M; ++m) w[((size_t)c * M + m) * N + i] = st[m] / total;
}

// KSDWeight._compute, ensembles/weights.py:396-441: per (cell, point) and model the IMQ kernel Stein
// discrepancy of the Ro observation samples against N(loc, scale) (scale = the member's VARIANCE, Q-SCALE):
//   g_a = -(x_a - loc) / scale^2 (:419);  k0(a, b) = the five terms of k_0_fun (:360-375) with dim = 1, c = 1,
//   beta = -1/2, q = 1 + (x_a - x_b)^2:  g_a g_b q^-1/2 + g_a d q^-3/2 - g_b d q^-3/2 + q^-3/2 - 3 q^-5/2 d^2;
//   ksd = sqrt(sum_ab k0) / Ro (:394);  weights = (1 / ksd) normalised over models (:434-438).
// The q powers are 1/sqrt(q) divided by q (once, twice) instead of three pow calls.
//
// Only g depends on the model; q, d and the three powers depend on the observation pair alone.  With
// xbar = mean_a x_a, u_a = x_a - xbar, delta = loc - xbar (so g_a = (delta - u_a) / scale^2), A = q^-1/2
// (symmetric), B = d q^-3/2 (antisymmetric, so sum_ab B_ab = 0) and C = q^-3/2 - 3 q^-5/2 d^2:
//   sum_ab k0 = (delta^2 SA - 2 delta SuA + SuuA) / scale^4 - 2 SuB / scale^2 + SC,
//   SA = sum_ab A_ab, SuA = sum_ab u_a A_ab, SuuA = sum_ab u_a u_b A_ab, SuB = sum_ab u_a B_ab = sum_{a<b} d^2 q^-3/2,
//   SC = sum_ab C_ab,
// i.e. Ro (Ro - 1) / 2 pair evaluations per POINT and a dozen flops per model instead of Ro^2 evaluations per
// (point, model): 45 ms -> see DESIGN.md at 4 M points x 24 models x 10 realisations.  Centring at xbar keeps the
// quadratic in delta as well conditioned as the direct double sum (worst relative error against a long-double
// evaluation: 1.1e-14 on climate-scale inputs against the oracle's 4.6e-15; 6e-10 for both on adversarial ones).
__global__ void k_ksd_weights(const double* __restrict__ loc, const double* __restrict__ scale,
                              const double* __restrict__ obs, int C, int M, int Ro, int N, double* __restrict__ w,
                              double* __restrict__ ksd_out, int smem_ok) {
    extern __shared__ double wstage[];
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)C * N) return;
    int c = (int)(gid / N), i = (int)(gid % N);
    WeightStage st(wstage, smem_ok, w + (size_t)c * M * N + i, (size_t)N);
    const double* ob = obs + (size_t)c * Ro * N + i;
    double xbar = 0.0;
    for (int a = 0; a < Ro; ++a) xbar += ob[(size_t)a * N];
    xbar /= (double)Ro;
    double SA = (double)Ro, SuA = 0.0, SuuA = 0.0, SuB = 0.0, SC = (double)Ro;  // the a == b terms: A = C = 1, B = 0
    for (int a = 0; a < Ro; ++a) {
        const double xa = ob[(size_t)a * N];
        const double ua = xa - xbar;
        SuA += ua;
        SuuA = fma(ua, ua, SuuA);
        for (int b = a + 1; b < Ro; ++b) {
            const double xb = ob[(size_t)b * N];
            const double ub = xb - xbar;
            const double d = xa - xb;
            const double d2 = d * d;
            const double q = 1.0 + d2;
            const double p05 = 1.0 / sqrt(q);  // q^-1/2
            const double p15 = p05 / q;        // q^-3/2
            const double p25 = p15 / q;        // q^-5/2
            SA += 2.0 * p05;
            SuA += (ua + ub) * p05;
            SuuA += 2.0 * (ua * ub) * p05;
            SuB += d2 * p15;
            SC += 2.0 * (p15 - 3.0 * p25 * d2);
        }
    }
    double total = 0.0;
    for (int m = 0; m < M; ++m) {
        size_t o = ((size_t)c * M + m) * N + i;
        const double l = loc[o], sc = scale[o];
        const double i2 = 1.0 / (sc * sc);
        const double dl = l - xbar;
        const double quad = (dl * dl) * SA - 2.0 * dl * SuA + SuuA;
        const double sum = (quad * i2 - 2.0 * SuB) * i2 + SC;
        const double ksd = sqrt(sum) / (double)Ro;
        if (ksd_out) ksd_out[o] = ksd;
        const double inv = 1.0 / ksd;
        st[m] = inv;
        if (inv == inv) total += inv;  // xarray .sum('model') skips NaN
    }
    for (int m = 0; m < M; ++m) w[((size_t)c * M + m) * N + i] = st[m] / total;
}

// nanmean over j of d[c, i, j, n], then normalise over i (weights.py:259,296,321 and :331).
// One thread per (cell, n); the M x M distances of a point are read once.
__global__ void k_w2_collapse(const double* __restrict__ d, int C, int M, int N, double* __restrict__ w, int smem_ok) {
    extern __shared__ double wstage[];
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)C * N) return;
    int c = (int)(gid / N), n = (int)(gid % N);
    WeightStage st(wstage, smem_ok, w + (size_t)c * M * N + n, (size_t)N);
    const double* dc = d + (size_t)c * M * M * N + n;
    double total = 0.0;
    for (int i = 0; i < M; ++i) {
        double s = 0.0, cnt = 0.0;
        for (int j = 0; j < M; ++j) {
            double v = dc[((size_t)i * M + j) * N];
            if (!isnan(v)) {
                s += v;
                cnt += 1.0;
            }
        }
        double m = s / cnt;
        st[i] = m;
        if (m == m) total += m;  // xarray .sum('model') skips NaN
    }
    for (int i = 0; i < M; ++i) w[((size_t)c * M + i) * N + n] = st[i] / total;
}

// mode="temporal" (weights.py:302-325): per point n and pair (i, j) the 1-dimensional
// full_cov=False W2 of wasserstein.py:36-45,  |m_i - m_j| + (v_i + v_j - 2 sqrt(sqrt(v_i) v_j sqrt(v_i))),
// v = the distributions' variance() (the caller passes it: variance**2 for the reference's
// dx.Normal(mean, variance)); nanmean over j; normalise over i.  Optionally writes the distances.
__global__ void k_similarity_pointwise(const double* __restrict__ mean, const double* __restrict__ var, int C, int M,
                                       int N, double* __restrict__ w, double* __restrict__ w2_out, int smem_ok) {
    extern __shared__ double wstage[];
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)C * N) return;
    int c = (int)(gid / N), n = (int)(gid % N);
    WeightStage st(wstage, smem_ok, w + (size_t)c * M * N + n, (size_t)N);
    const double* mc = mean + (size_t)c * M * N + n;
    const double* vc = var + (size_t)c * M * N + n;
    double total = 0.0;
    // When no pair distance is asked for and every mean and variance of the point is finite (variances >= 0), no
    // distance is NaN and sum_j dist_ij = sum_j |m_i - m_j| + (M v_i + sum_j v_j - 2 sqrt(v_i) sum_j sqrt(v_j))
    // (sqrt(r_i v_j r_i) = r_i r_j up to an ulp): M square roots per point instead of M^2, and a pair costs
    // a load, a subtraction and an addition.  Anything else takes the pair loop below, NaN-skipping as the
    // reference's nanmean does.
    if (!w2_out) {
        double Sv = 0.0, Sr = 0.0;
        bool clean = true;
        for (int j = 0; j < M; ++j) {
            const double mj = mc[(size_t)j * N], vj = vc[(size_t)j * N];
            const double rj = sqrt(vj);
            Sv += vj;
            Sr += rj;
            clean = clean && isfinite(mj) && isfinite(rj);
        }
        if (clean) {
            for (int i = 0; i < M; ++i) {
                const double mi = mc[(size_t)i * N], vi = vc[(size_t)i * N];
                double sabs = 0.0;
#pragma unroll 4
                for (int j = 0; j < M; ++j) sabs += fabs(mi - mc[(size_t)j * N]);
                const double m = (sabs + (((double)M * vi + Sv) - 2.0 * sqrt(vi) * Sr)) / (double)M;
                st[i] = m;
                if (m == m) total += m;  // xarray .sum('model') skips NaN
            }
            for (int i = 0; i < M; ++i) w[((size_t)c * M + i) * N + n] = st[i] / total;
            return;
        }
    }
    for (int i = 0; i < M; ++i) {
        const double mi = mc[(size_t)i * N], vi = vc[(size_t)i * N];
        const double ri = sqrt(vi);
        double s = 0.0, cnt = 0.0;
        for (int j = 0; j < M; ++j) {
            const double mj = mc[(size_t)j * N], vj = vc[(size_t)j * N];
            double dist = fabs(mi - mj) + ((vi + vj) - 2.0 * sqrt(ri * vj * ri));
            if (w2_out) w2_out[(((size_t)c * M + i) * M + j) * N + n] = dist;
            if (!isnan(dist)) {
                s += dist;
                cnt += 1.0;
            }
        }
        double m = s / cnt;
        st[i] = m;
        if (m == m) total += m;  // xarray .sum('model') skips NaN
    }
    for (int i = 0; i < M; ++i) w[((size_t)c * M + i) * N + n] = st[i] / total;
}

}  // namespace be
