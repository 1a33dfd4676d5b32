// SURVEY 8f "next" rows: CRPSWeight (ensembles/weights.py:444-515) and ModelSimilarityWeight
// (weights.py:214-333).  Per-point kernels over [C, M, N] arrays, HBM-bound.
#pragma once
#include "be_kernels.cuh"

namespace be {

constexpr double INV_SQRT_PI = 0.5641895835477563;

// A per-thread column of inputs that every model (or every pair) of the point re-reads: staged once in shared memory
// ([n][blockDim], conflict-free), or read in place when the CTA's share does not fit.  ncu r02x: with the WEIGHTS
// staged in shared memory (the round-1 layout) these kernels kept 6-9 CTAs x 24 KB resident, which left the L1 too
// small for the re-read inputs -- every re-read was an L2 round trip (long-scoreboard stalls 7-16 per issue, FP64
// pipe 25-40 % busy).  Now the inputs are staged and the un-normalised weights make one trip through global memory.
struct InStage {
    const double* p;
    size_t stride;
    __device__ __forceinline__ InStage(double* smem, bool in_smem, const double* src, size_t src_stride, int n) {
        if (in_smem) {
#pragma unroll 4
            for (int r = 0; r < n; ++r) smem[(size_t)r * blockDim.x + threadIdx.x] = src[(size_t)r * src_stride];
            p = smem + threadIdx.x;
            stride = blockDim.x;
        } else {
            p = src;
            stride = src_stride;
        }
    }
    __device__ __forceinline__ double operator[](int r) const { return p[(size_t)r * stride]; }
};
// w_m = st[m] / total over the models of one point: one reciprocal and M products when the total is an ordinary number
// (2^-1000 < total < 2^1000: within 1.5 ulp of the quotient); zero, subnormal, huge, infinite and NaN totals keep the
// divisions and with them the reference's 0/0 and x/inf patterns (as k_loglik_weights_mvn_tab does).
__device__ __forceinline__ void normalise_column(WeightStage& st, double* __restrict__ wp, size_t stride, int M, double total) {
    if (total > 0x1p-1000 && total < 0x1p1000) {
        const double inv = 1.0 / total;
#pragma unroll 4
        for (int m = 0; m < M; ++m) wp[(size_t)m * stride] = st[m] * inv;
    } else {
        for (int m = 0; m < M; ++m) wp[(size_t)m * stride] = st[m] / total;
    }
}
__host__ inline size_t in_stage_bytes(int n, int block) {
    size_t b = (size_t)n * block * sizeof(double);
    return b <= 64 * 1024 ? b : 0;
}

constexpr double INV_SQRT_2PI = 0.3989422804014327;

// properscoring.crps_gaussian: sig (z (2 Phi(z) - 1) + 2 phi(z) - 1/sqrt(pi)),  z = (x - mu) / sig.
// The bracket's first two terms are ONE even function of z,
//   G(z) = z erf(z / sqrt 2) + sqrt(2 / pi) exp(-z^2 / 2),   G' = erf(z / sqrt 2) =: E,   G'' = 2 phi,
//   G^(n) = 2 (-1)^n He_{n-2}(z) phi(z),
// so instead of an erf and an exp per evaluation (196 instructions, r01 ncu: the kernel was issue-bound at 7 % of the
// HBM peak) it is evaluated by a Taylor step from the nearest grid point c = k / 64 (|t| <= 1 / 128):
//   G(c + t) = G(c) + E(c) t + 2 phi(c) t^2 (1/2 - c t / 6 + He2 t^2 / 24 - He3 t^3 / 120 + He4 t^4 / 720),
// with G(c), E(c) from a 513-entry table in shared memory (crps_table.inc, correctly rounded by
// tools/make_crps_table.py), 2 phi(c) = G(c) - c E(c) (its cancellation error is scaled by t^2 / 2 < 3.1e-5) and the
// Hermite polynomials multiplied out (below): one 16-byte shared-memory load and ~20 FP64 operations.  Truncation
// 2 |He5 phi| t^7 / 5040 < 2e-18; worst relative error against a 60-digit evaluation 2.9e-16 (the same script,
// --check).  G(z) = |z| to the last bit from |z| = 8 on, which also covers the infinities; NaN propagates through t.
// The division by sig is a product with 1 / sig, formed once per model, when sig is an ordinary number
// (2^-500 < |sig| < 2^500; zero, subnormal, infinite and NaN scales keep the division and with it the
// reference's inf / NaN results).
#include "crps_table.inc"
constexpr int CRPS_TAB_N = 513;

// With u = c t and w = t^2 the bracket is A(u) + w B(u) + w^2 / 240, A = 1/2 - u/6 + u^2/24 - u^3/120 + u^4/720,
// B = -1/24 + u/40 - u^2/120 (the Hermite polynomials multiplied out): ten operations with constant coefficients.
// |z| and the test |z| >= 8 are done on the high word (integer pipe); 8.0 = 0x4020000000000000, and NaN passes the
// test and returns itself.
__device__ __forceinline__ double crps_kernel_G(double z, const double2* __restrict__ tab) {
    const double MAGIC = 6755399441055744.0;
    const int hi = __double2hiint(z) & 0x7fffffff;
    const double az = __hiloint2double(hi, __double2loint(z));
    const double v = fma(az, 64.0, MAGIC);  // k = rint(64 |z|) in the low word
    const unsigned k = min((unsigned)__double2loint(v), 512u);
    const double c = (v - MAGIC) * 0.015625;
    const double t = az - c;
    const double2 ge = tab[k];
    const double u = c * t, w = t * t;
    double a = fma(u, 1.0 / 720.0, -1.0 / 120.0);
    a = fma(a, u, 1.0 / 24.0);
    a = fma(a, u, -1.0 / 6.0);
    a = fma(a, u, 0.5);
    double b = fma(u, -1.0 / 120.0, 1.0 / 40.0);
    b = fma(b, u, -1.0 / 24.0);
    b = fma(w, 1.0 / 240.0, b);
    const double p = fma(w, b, a);
    const double p2 = fma(-c, ge.y, ge.x);  // 2 phi(c)
    const double g = fma(w, p2 * p, fma(t, ge.y, ge.x));
    return hi >= 0x40200000 ? az : g;
}

// one thread per (cell, point): weights.py:469-471 (mean over obs realisations), :507 (inverse),
// :510-511 (normalise over models).  scale = the distribution's stddev() -- for the dx.Normal the
// reference builds at :497 that is the member's VARIANCE (quirk Q-SCALE); the caller passes it.
// Dynamic shared memory: [513] double2 table | the point's observations [Ro][blockDim] (obs_smem).
__global__ void k_crps_weights(const double* __restrict__ loc, const double* __restrict__ scale,
                               const double* __restrict__ obs, int C, int M, int Ro, int N, double* __restrict__ w,
                               double* __restrict__ crps_mean, int obs_smem) {
    extern __shared__ __align__(16) double wstage_raw[];
    double2* tab = reinterpret_cast<double2*>(wstage_raw);
    for (int k = threadIdx.x; k < CRPS_TAB_N; k += blockDim.x) tab[k] = make_double2(CRPS_G[k], CRPS_E[k]);
    __syncthreads();
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)C * N) return;
    int c = (int)(gid / N), i = (int)(gid % N);
    WeightStage st(nullptr, false, w + (size_t)c * M * N + i, (size_t)N);
    const InStage ob(wstage_raw + 2 * (CRPS_TAB_N + 1), obs_smem, obs + (size_t)c * Ro * N + i, (size_t)N, Ro);
    double total = 0.0;
    const double inv_ro = 1.0 / (double)Ro, ro_c = (double)Ro * INV_SQRT_PI;
    size_t o = (size_t)c * M * N + i;
    double l_next = loc[o], sc_next = scale[o];
    for (int m = 0; m < M; ++m, o += N) {
        const double l = l_next, sc = sc_next;
        if (m + 1 < M) {  // the next model's pair is in flight while this one's Ro evaluations run
            l_next = loc[o + N];
            sc_next = scale[o + N];
        }
        const bool ordinary = fabs(sc) > 0x1p-500 && fabs(sc) < 0x1p500;
        const double inv_sc = 1.0 / sc;
        // sum_r sig (G(z_r) - 1/sqrt pi) = sig (sum_r G(z_r) - Ro / sqrt pi): the same infinities and NaN (an infinite
        // or NaN G makes either form inf / NaN, a zero sig makes z infinite or NaN), one operation per evaluation less
        double sg = 0.0;
        if (ordinary) {
#pragma unroll 2
            for (int r = 0; r < Ro; ++r) sg += crps_kernel_G((ob[r] - l) * inv_sc, tab);
        } else {
            for (int r = 0; r < Ro; ++r) sg += crps_kernel_G((ob[r] - l) / sc, tab);
        }
        double mean = sc * (sg - ro_c) * inv_ro;  // x * (1 / Ro): an ulp from x / Ro at most
        if (crps_mean) crps_mean[o] = mean;
        double inv = 1.0 / mean;
        st[m] = inv;
        if (inv == inv) total += inv;  // xarray .sum('model') skips NaN
    }
    normalise_column(st, w + (size_t)c * M * N + i, (size_t)N, M, total);
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
// 1 / sqrt(q) for q >= 1 (finite): MUFU seed (~2^-22) and one third-order step y (1 + e / 2 + 3 e^2 / 8),
// e = 1 - q y^2 -- the last-bit error of a correctly rounded 1 / sqrt(q) or one ulp more, against the ~30
// instructions of sqrt followed by a division.  Non-finite q gives NaN (the reference's sums are NaN there too:
// an infinite observation makes d^2 q^-3/2 = inf * 0).
__device__ __forceinline__ double rsqrt_ge1(double q) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
    const double e = fma(-q * y, y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}

__global__ void k_ksd_weights(const double* __restrict__ loc, const double* __restrict__ scale,
                              const double* __restrict__ obs, int C, int M, int Ro, int N, double* __restrict__ w,
                              double* __restrict__ ksd_out, int obs_smem) {
    extern __shared__ double wstage[];
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)C * N) return;
    int c = (int)(gid / N), i = (int)(gid % N);
    WeightStage st(nullptr, false, w + (size_t)c * M * N + i, (size_t)N);
    const InStage ob(wstage, obs_smem, obs + (size_t)c * Ro * N + i, (size_t)N, Ro);
    double xbar = 0.0;
    for (int a = 0; a < Ro; ++a) xbar += ob[a];
    xbar /= (double)Ro;
    double SA = (double)Ro, SuA = 0.0, SuuA = 0.0, SuB = 0.0, SC = (double)Ro;  // the a == b terms: A = C = 1, B = 0
    for (int a = 0; a < Ro; ++a) {
        const double xa = ob[a];
        const double ua = xa - xbar;
        SuA += ua;
        SuuA = fma(ua, ua, SuuA);
        for (int b = a + 1; b < Ro; ++b) {
            const double xb = ob[b];
            const double ub = xb - xbar;
            const double d = xa - xb;
            const double d2 = d * d;
            const double q = 1.0 + d2;
            const double p05 = rsqrt_ge1(q);   // q^-1/2
            const double qi = p05 * p05;       // q^-1 (q >= 1: no division needed, two ulps)
            const double p15 = p05 * qi;       // q^-3/2
            const double p25 = p15 * qi;       // q^-5/2
            SA += 2.0 * p05;
            SuA += (ua + ub) * p05;
            SuuA += 2.0 * (ua * ub) * p05;
            SuB += d2 * p15;
            SC += 2.0 * (p15 - 3.0 * p25 * d2);
        }
    }
    double total = 0.0;
#pragma unroll 4
    for (int m = 0; m < M; ++m) {  // unrolled so that four models' loads are in flight together
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
    normalise_column(st, w + (size_t)c * M * N + i, (size_t)N, M, total);
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
    normalise_column(st, w + (size_t)c * M * N + n, (size_t)N, M, total);
}

// mode="temporal" (weights.py:302-325): per point n and pair (i, j) the 1-dimensional
// full_cov=False W2 of wasserstein.py:36-45,  |m_i - m_j| + (v_i + v_j - 2 sqrt(sqrt(v_i) v_j sqrt(v_i))),
// v = the distributions' variance() (the caller passes it: variance**2 for the reference's
// dx.Normal(mean, variance)); nanmean over j; normalise over i.  Optionally writes the distances.
__global__ void k_similarity_pointwise(const double* __restrict__ mean, const double* __restrict__ var, int C, int M,
                                       int N, double* __restrict__ w, double* __restrict__ w2_out, int mean_smem) {
    extern __shared__ double wstage[];
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)C * N) return;
    int c = (int)(gid / N), n = (int)(gid % N);
    WeightStage st(nullptr, false, w + (size_t)c * M * N + n, (size_t)N);
    const InStage ms(wstage, mean_smem, mean + (size_t)c * M * N + n, (size_t)N, M);  // the M means: read M + 4 times each
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
#pragma unroll 4
        for (int j = 0; j < M; ++j) {
            const double mj = ms[j], vj = vc[(size_t)j * N];
            const double rj = sqrt(vj);
            Sv += vj;
            Sr += rj;
            clean = clean && isfinite(mj) && isfinite(rj);
        }
        if (clean) {
            // four rows i at a time against one load of m_j: the pair loop was bound by its M^2 L1 loads per point
            // (same summation order per row, so the sums are bit-identical to the one-row form)
            for (int i0 = 0; i0 < M; i0 += 4) {
                double mi[4], sabs[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int u = 0; u < 4; ++u) mi[u] = ms[min(i0 + u, M - 1)];
                double vi4[4];  // loaded ahead of the pair loop that hides their latency
#pragma unroll
                for (int u = 0; u < 4; ++u) vi4[u] = vc[(size_t)min(i0 + u, M - 1) * N];
#pragma unroll 2
                for (int j = 0; j < M; ++j) {
                    const double mj = ms[j];
#pragma unroll
                    for (int u = 0; u < 4; ++u) sabs[u] += fabs(mi[u] - mj);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u;
                    if (i >= M) break;
                    const double vi = vi4[u];
                    const double m = (sabs[u] + (((double)M * vi + Sv) - 2.0 * sqrt(vi) * Sr)) / (double)M;
                    st[i] = m;
                    if (m == m) total += m;  // xarray .sum('model') skips NaN
                }
            }
            normalise_column(st, w + (size_t)c * M * N + n, (size_t)N, M, total);
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
    normalise_column(st, w + (size_t)c * M * N + n, (size_t)N, M, total);
}

}  // namespace be
