// DTW barycentre averaging on the device (the step that produces y_mean,
// ensembles/models.py:176-178 via tslearn, and the reference's own ensembles/dtwa.py).
//
// The DTW table cost[i,j] = (a_i - x_j)^2 + min(cost[i-1,j-1], cost[i-1,j], cost[i,j-1]) is filled as a
// systolic wavefront: thread t owns the W columns [tW, tW+W) and works on row i = s - t at step s,
// so the only communication is "my last column of this row" to thread t+1 (a warp shuffle; one
// shared-memory slot per warp boundary and one barrier per step when a pair needs more than a
// warp).  Rows of the previous step live in registers.  Each thread records its W 2-bit argmin
// codes per step in ONE word, laid out step-major ([step][thread]) so that a step's stores are
// coalesced; the backtrack kernel maps (i, j) -> word (i + j/W, j/W).
//
// Every product and sum is an explicit round-to-nearest intrinsic (no FMA contraction): the
// path is a discrete function of the table, so the table must round exactly as the CPU oracle
// (oracle/dba.c, -ffp-contract=off) and the NumPy/numba code it restates.
#pragma once
#include <stdint.h>

namespace be {

constexpr int DTW_TIE_TSLEARN = 0;  // argmin([diag, top, left]), first minimum wins (tslearn _return_path)
constexpr int DTW_TIE_DTWA = 1;     // ensembles/dtwa.py:113-129

template <int W> struct DtwWord { typedef uint32_t type; };
template <> struct DtwWord<1> { typedef uint8_t type; };
template <> struct DtwWord<2> { typedef uint8_t type; };
template <> struct DtwWord<4> { typedef uint8_t type; };
template <> struct DtwWord<8> { typedef uint16_t type; };

__host__ __device__ inline int dtw_threads_used(int T, int W) { return (T + W - 1) / W; }
__host__ __device__ inline int dtw_steps(int T, int W) { return T + dtw_threads_used(T, W) - 1; }

// One row of the table for the W columns a thread owns: up[] (the previous row) is replaced by this row, left
// enters as cost[i, j0-1] and leaves as cost[i, j0+W-1], diag enters as cost[i-1, j0-1]; returns the W 2-bit
// argmin codes (0 diag, 1 top = (i-1, j), 2 left = (i, j-1)).
template <int W, int TIE>
__device__ __forceinline__ unsigned dtw_row(double ai, const double (&xj)[W], double (&up)[W], double& left,
                                            double diag) {
    unsigned codes = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        const double diff = __dsub_rn(ai, xj[k]);
        const double d = __dmul_rn(diff, diff);
        const double top = up[k];
        // Branch-free: under either tie rule the VALUE is min(diag, top, left) (equal candidates are the same double:
        // the table holds no NaN and no negative zero), only the 2-bit code depends on the rule.  Written with
        // nested if / else the dtwa.py rule compiled to a divergent branch per cell (BSSY / BRA / BSYNC, both arms
        // executed whenever the lanes of a warp disagreed); the tslearn rule was already predicated, and its time did
        // not change (cfg2: 265 ms per 50 iterations before and after, r02K).  Two compare-and-select pairs, not fmin:
        // sm_100a has no FP64 min instruction, fmin() expands to ~8 instructions with its NaN handling.
        const bool top_lt_diag = top < diag;
        const double m1 = top_lt_diag ? top : diag;
        const bool left_lt_m1 = left < m1;
        const double m = left_lt_m1 ? left : m1;
        unsigned code;
        if (TIE == DTW_TIE_TSLEARN) {  // argmin([diag, top, left]), first minimum wins
            code = left_lt_m1 ? 2u : (top_lt_diag ? 1u : 0u);
        } else {                       // ensembles/dtwa.py:113-129 (diag <= top is !(top < diag): no NaN in the table)
            const unsigned c_diag = top_lt_diag ? 1u : 0u, c_left = left <= top ? 2u : 1u;
            code = diag <= left ? c_diag : c_left;
        }
        const double cur = __dadd_rn(m, d);
        diag = top;
        up[k] = cur;
        left = cur;
        codes |= code << (2 * k);
    }
    return codes;
}

// resident CTAs per SM the DP kernel is compiled for: the kernel is issue-bound, so what matters is that
// the register budget leaves no spill and that the CTAs of a launch fill the SMs without a long tail
// (T = 3012: 14 columns x 7 warps -> 94 registers x 224 threads -> 3 CTAs per SM)
// (W = 14 x 7 warps at 3 CTAs per SM spills nine of the x_j to local memory -- 80 registers; at 2 CTAs per SM nothing
// spills -- 110 registers -- and the kernel measured 2 % faster at cfg2, equal at cfg3 / cfg4: r02K.  3 stays.)
#ifndef BE_DTW_CTAS_W14
#define BE_DTW_CTAS_W14 3
#endif
__host__ __device__ constexpr int dtw_min_ctas(int W, int NWARPS) {
    return NWARPS == 1 ? 4 : NWARPS == 7 ? BE_DTW_CTAS_W14 : W <= 8 ? 3 : 2;
}

// (A barrier-free variant was tried for the multi-warp form: neighbouring warps decoupled through a 64-row
// shared-memory ring, flag and value in one 8-byte word, consumer kept 24 rows behind.  Bit-exact, but 2.5x
// SLOWER -- 9.5 G instead of 3.8 G warp instructions per launch, a quarter of the stall samples at the
// shuffle that re-gathers the lanes after the first lane's divergent wait (ncu r01o) -- so the barrier stays.)
// One pair = (row sequence a = A[pair / R], column sequence x = X[(pair / x_group) * R + pair % R]).
// NWARPS == 1: one warp per pair, 4 pairs per 128-thread CTA, shuffles only.
// NWARPS  > 1: one CTA of NWARPS warps per pair.
template <int W, int NWARPS, int TIE, bool DIRS>
__global__ void __launch_bounds__(NWARPS == 1 ? 128 : NWARPS * 32, dtw_min_ctas(W, NWARPS))
k_dtw_dp(const double* __restrict__ A, const double* __restrict__ X, int T, int R, int x_group, int n_pairs,
         const int* __restrict__ active, typename DtwWord<W>::type* __restrict__ dirs, size_t dirs_stride,
         double* __restrict__ sqcost) {
    typedef typename DtwWord<W>::type word_t;
    constexpr int NT = NWARPS * 32;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int pair, t;
    if (NWARPS == 1) {
        pair = blockIdx.x * 4 + warp;
        t = lane;
        if (pair >= n_pairs) return;
    } else {
        pair = blockIdx.x;
        t = threadIdx.x;
    }
    if (active && !active[pair / R]) return;  // uniform over the threads that share barriers
    const double* a = A + (size_t)(pair / R) * T;
    const double* x = X + ((size_t)(pair / x_group) * R + pair % R) * T;
    const int nthr = dtw_threads_used(T, W);
    const int steps = T + nthr - 1;
    const int j0 = t * W;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double xj[W], up[W];
#pragma unroll
    for (int k = 0; k < W; ++k) {
        xj[k] = (j0 + k < T) ? x[j0 + k] : 0.0;
        up[k] = INF;
    }
    double diag_in = (t == 0) ? 0.0 : INF;  // cost[-1,-1] = 0 (tslearn's border; dtwa.py:53 cost[0,0] = delta[0,0])
    double out = INF;
    double a_cur = (t == 0) ? a[0] : 0.0;
    word_t* drow = DIRS ? dirs + (size_t)pair * dirs_stride + t : nullptr;
    __shared__ double edge[2][NWARPS > 1 ? NWARPS : 1];
    for (int s = 0; s < steps; ++s) {
        double in = __shfl_up_sync(0xffffffffu, out, 1);
        if (NWARPS > 1) {
            if (lane == 0 && warp > 0) in = edge[(s & 1) ^ 1][warp - 1];
        }
        if (t == 0) in = INF;
        const int i = s - t;
        if (i >= 0 && i < T && t < nthr) {
            double left = in;
            const unsigned codes = dtw_row<W, TIE>(a_cur, xj, up, left, diag_in);
            out = left;
            diag_in = in;
            if (DIRS) drow[(size_t)s * NT] = (word_t)codes;
            if (i == T - 1 && t == nthr - 1) {
                const int kl = (T - 1) - j0;
                double fin = 0.0;
#pragma unroll
                for (int k = 0; k < W; ++k)
                    if (k == kl) fin = up[k];
                sqcost[pair] = fin;
            }
        }
        const int in1 = i + 1;
        a_cur = (in1 >= 0 && in1 < T) ? __ldg(a + in1) : 0.0;
        if (NWARPS > 1) {
            if (lane == 31) edge[s & 1][warp] = out;
            __syncthreads();
        }
    }
}

// Path walk for the one-warp-per-pair shapes (T <= 512).  In the step-major layout the words of 32 consecutive
// steps of a pair are one contiguous block (32 steps x 32 lanes x 1/2/4 bytes), so the warp copies such a
// block to shared memory with fully used 128-byte lines and walks inside it -- a move lowers the step index
// i + j/W by at most two, so a block lasts for at least 16 moves and typically ~40.  With the block copy
// batched and x in a register window the kernel is ISSUE-bound (ncu r01o: issue active 80 %, ~50 warp
// instructions per move, every lane walking the same path); a lane-per-pair walk would use the lanes but
// turn every move into a dependent L2 round trip, which only pays with >= 10^4 pairs in flight -- not done.
// (A fused fill+walk kernel with the words of the whole pair in shared memory was slower: 18 KB per pair
// leaves 12 warps per SM for the issue-bound fill.)
template <int W>
__global__ void __launch_bounds__(128)
k_dtw_backtrack_w1(const double* __restrict__ X, int T, int R, int n_pairs, const int* __restrict__ active,
                   const typename DtwWord<W>::type* __restrict__ dirs, size_t dirs_stride, double* __restrict__ v,
                   double* __restrict__ wx) {
    typedef typename DtwWord<W>::type word_t;
    __shared__ __align__(16) word_t win[4][32 * 32];
    const int lane = threadIdx.x & 31;
    const int wslot = threadIdx.x >> 5;
    const int pair = blockIdx.x * 4 + wslot;
    if (pair >= n_pairs) return;
    if (active && !active[pair / R]) return;
    const double* x = X + (size_t)pair * T;
    const word_t* d = dirs + (size_t)pair * dirs_stride;
    double* vo = v + (size_t)pair * T;
    double* wo = wx + (size_t)pair * T;
    word_t* wn = win[wslot];
    int i = T - 1, j = T - 1;
    double acc = 0.0, cnt = 0.0;
    // x[j] comes from a 32-wide register window (lane l holds x[jw - l]) refilled by one coalesced load: a
    // dependent global load per move would stall the in-order warp for an L2 round trip every step
    int jw = j;
    double xw = (jw - lane >= 0) ? x[jw - lane] : 0.0;
    for (;;) {
        const int s_hi = i + j / W;            // step index of the current cell
        const int s_lo = max(s_hi - 31, 0);    // block = steps s_lo .. s_hi
        __syncwarp();
        {   // all of a lane's 16-byte pieces are requested before the first is stored: one memory latency per block
            constexpr int PIECES = 2 * (int)sizeof(word_t);  // 32 steps x 32 lanes x sizeof(word) / 16 B / 32 lanes
            const int n_pieces = (s_hi - s_lo + 1) * 2 * (int)sizeof(word_t);
            const uint4* src = reinterpret_cast<const uint4*>(d + (size_t)s_lo * 32);
            uint4* dst = reinterpret_cast<uint4*>(wn);
            uint4 tmp[PIECES];
#pragma unroll
            for (int k = 0; k < PIECES; ++k)
                if (k * 32 + lane < n_pieces) tmp[k] = src[k * 32 + lane];
#pragma unroll
            for (int k = 0; k < PIECES; ++k)
                if (k * 32 + lane < n_pieces) dst[k * 32 + lane] = tmp[k];
        }
        __syncwarp();
        bool done = false;
        for (;;) {
            if (j < jw - 31) {
                jw = j;
                xw = (jw - lane >= 0) ? x[jw - lane] : 0.0;
            }
            cnt += 1.0;
            acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, xw, jw - j));
            if (i == 0 && j == 0) { done = true; break; }
            const int t = j / W;
            const unsigned code = ((unsigned)wn[(i + t - s_lo) * 32 + t] >> (2 * (j - t * W))) & 3u;
            int ni = i - (code != 2u), nj = j - (code != 1u);
            // a NaN table (NaN inputs) can record "diag" on a border; keep the walk inside the table
            if (ni < 0) { ni = 0; nj = j - 1; }
            if (nj < 0) { nj = 0; ni = i - 1; }
            if (ni != i) {
                if (lane == 0) { vo[i] = cnt; wo[i] = acc; }
                cnt = 0.0;
                acc = 0.0;
            }
            i = ni;
            j = nj;
            if (i + j / W < s_lo) break;
        }
        if (done) {
            if (lane == 0) { vo[0] = cnt; wo[0] = acc; }
            break;
        }
    }
}

// Path walk with ONE THREAD per pair, for launches with many pairs (>= 8192: a grid of cells).  The warp-per-
// pair walks above and below are issue-bound with every lane repeating the same ~50 instructions per move;
// here the 32 lanes walk 32 different paths, and what a move costs is one dependent 1-4 byte load of its
// argmin word (an L2 / HBM round trip) -- paid once for thousands of walkers in flight.  x[j-1] is fetched
// together with the word, so that the next cell's x never adds a second round trip.  Same accumulation order.
template <int W>
__global__ void __launch_bounds__(128)
k_dtw_backtrack_lane(const double* __restrict__ X, int T, int R, int n_pairs, const int* __restrict__ active,
                     const typename DtwWord<W>::type* __restrict__ dirs, size_t dirs_stride, double* __restrict__ v,
                     double* __restrict__ wx) {
    typedef typename DtwWord<W>::type word_t;
    const int pair = blockIdx.x * 128 + threadIdx.x;
    if (pair >= n_pairs) return;
    if (active && !active[pair / R]) return;
    const double* x = X + (size_t)pair * T;
    const word_t* d = dirs + (size_t)pair * dirs_stride;
    double* vo = v + (size_t)pair * T;
    double* wo = wx + (size_t)pair * T;
    int i = T - 1, j = T - 1;
    double acc = 0.0, cnt = 0.0;
    double xv = x[j];
    for (;;) {
        cnt += 1.0;
        acc = __dadd_rn(acc, xv);
        if (i == 0 && j == 0) break;
        const int t = j / W;
        const unsigned word = d[(size_t)(i + t) * 32 + t];
        const double xm1 = x[j > 0 ? j - 1 : 0];
        const unsigned code = (word >> (2 * (j - t * W))) & 3u;
        int ni = i - (code != 2u), nj = j - (code != 1u);
        if (ni < 0) { ni = 0; nj = j - 1; }
        if (nj < 0) { nj = 0; ni = i - 1; }
        if (nj != j) xv = xm1;
        if (ni != i) {
            vo[i] = cnt;
            wo[i] = acc;
            cnt = 0.0;
            acc = 0.0;
        }
        i = ni;
        j = nj;
    }
    vo[0] = cnt;
    wo[0] = acc;
}

// Walks the optimal path of one pair back from (T-1, T-1) (tslearn _return_path; dtwa.py:131-139).
// One warp per pair: the lanes fetch the direction words of 32 rows x 2 column strips around the
// current cell in one coalesced-ish gather, then the warp walks inside that window with shuffles.
// v[pair, i] = number of path cells in row i, wx[pair, i] = sum of x_j over them in walk order
// (j decreasing) -- the summation order oracle/dba.c defines.
template <int W, int NWARPS>
__global__ void __launch_bounds__(128)
k_dtw_backtrack(const double* __restrict__ X, int T, int R, int n_pairs, const int* __restrict__ active,
                const typename DtwWord<W>::type* __restrict__ dirs, size_t dirs_stride, double* __restrict__ v,
                double* __restrict__ wx) {
    typedef typename DtwWord<W>::type word_t;
    constexpr int NT = NWARPS * 32;
    const int lane = threadIdx.x & 31;
    const int pair = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (pair >= n_pairs) return;
    if (active && !active[pair / R]) return;
    const double* x = X + (size_t)pair * T;
    const word_t* d = dirs + (size_t)pair * dirs_stride;
    double* vo = v + (size_t)pair * T;
    double* wo = wx + (size_t)pair * T;
    int i = T - 1, j = T - 1;
    double acc = 0.0, cnt = 0.0;
    int jw = j;  // 32-wide register window of x (see k_dtw_backtrack_w1)
    double xw = (jw - lane >= 0) ? x[jw - lane] : 0.0;
    for (;;) {
        const int i0 = i, t0 = j / W;
        const int row = i0 - lane;
        unsigned wA = 0, wB = 0;
        if (row >= 0) {
            wA = d[(size_t)(row + t0) * NT + t0];
            if (t0 > 0) wB = d[(size_t)(row + t0 - 1) * NT + (t0 - 1)];
        }
        bool done = false;
        for (;;) {
            if (j < jw - 31) {
                jw = j;
                xw = (jw - lane >= 0) ? x[jw - lane] : 0.0;
            }
            cnt += 1.0;
            acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, xw, jw - j));
            if (i == 0 && j == 0) { done = true; break; }
            const int t = j / W;
            const unsigned word = __shfl_sync(0xffffffffu, (t == t0) ? wA : wB, i0 - i);
            const unsigned code = (word >> (2 * (j - t * W))) & 3u;
            int ni = i - (code != 2u), nj = j - (code != 1u);
            // a NaN table (NaN inputs) can record "diag" on a border; keep the walk inside the table
            if (ni < 0) { ni = 0; nj = j - 1; }
            if (nj < 0) { nj = 0; ni = i - 1; }
            if (ni != i) {
                if (lane == 0) { vo[i] = cnt; wo[i] = acc; }
                cnt = 0.0;
                acc = 0.0;
            }
            i = ni;
            j = nj;
            if (i0 - i >= 32 || j / W < t0 - 1) break;
        }
        if (done) {
            if (lane == 0) { vo[0] = cnt; wo[0] = acc; }
            break;
        }
    }
}

// tslearn _subgradient_update_barycenter + the loop tail of dtw_barycenter_averaging_subgradient:
//   delta = sum_k (v_k * c - wx_k)  (k increasing, "+= v_k c" then "-= wx_k");  c -= (2 eta / R) delta;
//   cost = sum_k sqrt(sq_k)^2 / R;  |cost_prev - cost| < tol -> stop;  cost_prev < cost -> keep cost_prev
//   (tslearn only warns);  else cost_prev = cost.
// One CTA per problem. state[b]: active flag; n_active: device counter the host polls.
__global__ void k_dba_subgradient_update(double* __restrict__ bary, const double* __restrict__ v,
                                         const double* __restrict__ wx, const double* __restrict__ sqcost, int T, int R,
                                         double step, double tol, int* __restrict__ active,
                                         double* __restrict__ cost_prev, double* __restrict__ cost_last,
                                         int* __restrict__ n_iter, int* __restrict__ n_active) {
    const int b = blockIdx.x;
    if (!active[b]) return;
    double* c = bary + (size_t)b * T;
    const double* vb = v + (size_t)b * R * T;
    const double* wb = wx + (size_t)b * R * T;
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const double ci = c[i];
        double delta = 0.0;
        for (int k = 0; k < R; ++k) {
            delta = __dadd_rn(delta, __dmul_rn(vb[(size_t)k * T + i], ci));
            delta = __dsub_rn(delta, wb[(size_t)k * T + i]);
        }
        c[i] = __dsub_rn(ci, __dmul_rn(step, delta));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double cost = 0.0;
        for (int k = 0; k < R; ++k) {
            const double dist = sqrt(sqcost[(size_t)b * R + k]);
            cost = __dadd_rn(cost, __dmul_rn(dist, dist));
        }
        cost = cost / (double)R;
        n_iter[b] += 1;
        cost_last[b] = cost;
        const double prev = cost_prev[b];
        if (fabs(prev - cost) < tol) {
            active[b] = 0;
            atomicSub(n_active, 1);
        } else if (prev < cost) {
        } else {
            cost_prev[b] = cost;
        }
    }
}

// ensembles/dtwa.py:23-37: ss[c] = sum_k squared_DTW(series_c, series_k) (sequential from 0), first
// strict minimum wins; centre := series[medoid] (dtwa.py:15).  sq [B, R, R].  One CTA per problem.
__global__ void k_dba_medoid(const double* __restrict__ X, const double* __restrict__ sq, int T, int R,
                             double* __restrict__ center, int* __restrict__ medoid) {
    const int b = blockIdx.x;
    __shared__ int best_c;
    if (threadIdx.x == 0) {
        int m = -1;
        double best = 1e20;
        for (int c = 0; c < R; ++c) {
            double ss = 0.0;
            for (int k = 0; k < R; ++k) ss = __dadd_rn(ss, sq[((size_t)b * R + c) * R + k]);
            if (m == -1 || ss < best) { best = ss; m = c; }
        }
        best_c = m;
        if (medoid) medoid[b] = m;
    }
    __syncthreads();
    const double* src = X + ((size_t)b * R + best_c) * T;
    for (int i = threadIdx.x; i < T; i += blockDim.x) center[(size_t)b * T + i] = src[i];
}

// ensembles/dtwa.py:141: centre = updated_center / n_elements (sums grouped per series, k increasing)
__global__ void k_dba_mean_update(double* __restrict__ center, const double* __restrict__ v,
                                  const double* __restrict__ wx, int B, int T, int R) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)B * T) return;
    const int b = (int)(gid / T), i = (int)(gid % T);
    double s = 0.0, n = 0.0;
    for (int k = 0; k < R; ++k) {
        s = __dadd_rn(s, wx[((size_t)b * R + k) * T + i]);
        n = __dadd_rn(n, v[((size_t)b * R + k) * T + i]);
    }
    center[gid] = s / n;
}

// _init_avg (tslearn dba.py): the mean over series when barycenter_size == T
__global__ void k_dba_init_mean(const double* __restrict__ X, int B, int R, int T, double* __restrict__ bary) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)B * T) return;
    const int b = (int)(gid / T), i = (int)(gid % T);
    double s = 0.0;
    for (int k = 0; k < R; ++k) s = __dadd_rn(s, X[((size_t)b * R + k) * T + i]);
    bary[gid] = s / (double)R;
}

__global__ void k_dba_state_init(int B, int* active, double* cost_prev, double* cost_last, int* n_iter, int* n_active) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) *n_active = B;
    if (b >= B) return;
    active[b] = 1;
    cost_prev[b] = __longlong_as_double(0x7ff0000000000000LL);
    cost_last[b] = __longlong_as_double(0x7ff0000000000000LL);
    n_iter[b] = 0;
}

}  // namespace be
