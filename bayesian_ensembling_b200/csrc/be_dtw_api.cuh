// C ABI, third part (included by be_api.cu): DTW barycentre averaging.
// Host orchestration over dtw_kernels.cuh.
#pragma once

namespace {

constexpr int DTW_LANE_WALK_MIN_PAIRS = 8192;  // thread-per-pair path walk from this many pairs on (T <= 512)

struct DtwShape {
    int W, NW;        // columns per thread, warps per pair
    int word_bytes;   // bytes of one direction word
};

// T -> (W, NWARPS): one warp per pair while T <= 512, else one CTA of 7-8 warps per pair
inline bool dtw_shape(int T, DtwShape& s) {
    if (T <= 32) s = {1, 1, 1};
    else if (T <= 64) s = {2, 1, 1};
    else if (T <= 128) s = {4, 1, 1};
    else if (T <= 256) s = {8, 1, 2};
    else if (T <= 512) s = {16, 1, 4};
    else if (T <= 1024) s = {4, 8, 1};
    else if (T <= 2048) s = {8, 8, 2};
    else if (T <= 3136) s = {14, 7, 4};
    else if (T <= 4096) s = {16, 8, 4};
    else return false;
    return true;
}

inline size_t dtw_dirs_stride(int T, const DtwShape& s) {  // words per pair
    return (size_t)dtw_steps(T, s.W) * (size_t)(s.NW * 32);
}

struct DtwBuffers {
    void* dirs;
    double *v, *wx, *sq, *cost_prev, *cost_last_own;
    int *active, *n_iter_own, *n_active;
};

size_t dtw_core_bytes(int B, int R, int T) {
    DtwShape s;
    if (!dtw_shape(T, s)) return 0;
    const size_t pairs = (size_t)B * R;
    return align_up(pairs * dtw_dirs_stride(T, s) * s.word_bytes, 256) + 2 * align_up(pairs * T * 8, 256) +
           align_up(pairs * R * 8, 256) + 2 * align_up((size_t)B * 8, 256) + 2 * align_up((size_t)B * 4, 256) + 256;
}

bool carve_dtw(Carver& cv, int B, int R, int T, const DtwShape& s, DtwBuffers& w) {
    const size_t pairs = (size_t)B * R;
    w.dirs = cv.take<char>(pairs * dtw_dirs_stride(T, s) * s.word_bytes);
    w.v = cv.take<double>(pairs * T);
    w.wx = cv.take<double>(pairs * T);
    w.sq = cv.take<double>(pairs * R);
    w.cost_prev = cv.take<double>(B);
    w.cost_last_own = cv.take<double>(B);
    w.active = cv.take<int>(B);
    w.n_iter_own = cv.take<int>(B);
    w.n_active = cv.take<int>(1);
    return w.n_active != nullptr;
}

template <int W, int NW, int TIE, bool DIRS>
int launch_dtw_dp_t(be_ctx* ctx, const double* A, const double* X, int T, int R, int x_group, int n_pairs,
                    const int* active, void* dirs, size_t stride, double* sq) {
    typedef typename DtwWord<W>::type word_t;
    // ALGORITHMIC work per launch: T^2 cells per pair, 5 fp64 operations each (sub, mul, two compares, add);
    // bytes: the 2-bit argmin code per cell (if recorded) + the two sequences
    Prof p(ctx, F_DTW_DP, 5.0 * T * (double)T * n_pairs, ((DIRS ? 0.25 : 0.0) * T * (double)T + 16.0 * T) * n_pairs);
    const unsigned grid = NW == 1 ? (unsigned)((n_pairs + 3) / 4) : (unsigned)n_pairs;
    k_dtw_dp<W, NW, TIE, DIRS><<<grid, NW == 1 ? 128 : NW * 32, 0, ctx->stream>>>(A, X, T, R, x_group, n_pairs, active,
                                                                                 (word_t*)dirs, stride, sq);
    BE_LAUNCHED();
    return BE_OK;
}

template <int W, int NW>
int launch_dtw_dp_w(be_ctx* ctx, int tie, bool want_dirs, const double* A, const double* X, int T, int R, int x_group,
                    int n_pairs, const int* active, void* dirs, size_t stride, double* sq) {
    if (want_dirs && tie == DTW_TIE_TSLEARN)
        return launch_dtw_dp_t<W, NW, DTW_TIE_TSLEARN, true>(ctx, A, X, T, R, x_group, n_pairs, active, dirs, stride, sq);
    if (want_dirs) return launch_dtw_dp_t<W, NW, DTW_TIE_DTWA, true>(ctx, A, X, T, R, x_group, n_pairs, active, dirs, stride, sq);
    // the table's VALUES do not depend on the tie rule: one cost-only instantiation serves both
    return launch_dtw_dp_t<W, NW, DTW_TIE_DTWA, false>(ctx, A, X, T, R, x_group, n_pairs, active, dirs, stride, sq);
}

#define BE_DTW_DISPATCH(FN, ...)                                          \
    do {                                                                  \
        switch (shape.NW * 100 + shape.W) {                               \
            case 101: return FN<1, 1>(__VA_ARGS__);                       \
            case 102: return FN<2, 1>(__VA_ARGS__);                       \
            case 104: return FN<4, 1>(__VA_ARGS__);                       \
            case 108: return FN<8, 1>(__VA_ARGS__);                       \
            case 116: return FN<16, 1>(__VA_ARGS__);                      \
            case 804: return FN<4, 8>(__VA_ARGS__);                       \
            case 808: return FN<8, 8>(__VA_ARGS__);                       \
            case 714: return FN<14, 7>(__VA_ARGS__);                      \
            case 816: return FN<16, 8>(__VA_ARGS__);                      \
        }                                                                 \
        return BE_ERR_UNSUPPORTED;                                        \
    } while (0)

int launch_dtw_dp(be_ctx* ctx, const DtwShape& shape, int tie, bool want_dirs, const double* A, const double* X, int T,
                  int R, int x_group, int n_pairs, const int* active, void* dirs, size_t stride, double* sq) {
    BE_DTW_DISPATCH(launch_dtw_dp_w, ctx, tie, want_dirs, A, X, T, R, x_group, n_pairs, active, dirs, stride, sq);
}

template <int W, int NW>
int launch_dtw_backtrack_w(be_ctx* ctx, const double* X, int T, int R, int n_pairs, const int* active, const void* dirs,
                           size_t stride, double* v, double* wx) {
    typedef typename DtwWord<W>::type word_t;
    Prof p(ctx, F_DTW_BACK, 0.0, 40.0 * T * n_pairs);
    if (NW == 1 && n_pairs >= DTW_LANE_WALK_MIN_PAIRS)
        k_dtw_backtrack_lane<W><<<(unsigned)((n_pairs + 127) / 128), 128, 0, ctx->stream>>>(
            X, T, R, n_pairs, active, (const word_t*)dirs, stride, v, wx);
    else if (NW == 1)
        k_dtw_backtrack_w1<W><<<(unsigned)((n_pairs + 3) / 4), 128, 0, ctx->stream>>>(X, T, R, n_pairs, active,
                                                                                     (const word_t*)dirs, stride, v, wx);
    else
        k_dtw_backtrack<W, NW><<<(unsigned)((n_pairs + 3) / 4), 128, 0, ctx->stream>>>(
            X, T, R, n_pairs, active, (const word_t*)dirs, stride, v, wx);
    BE_LAUNCHED();
    return BE_OK;
}

int launch_dtw_backtrack(be_ctx* ctx, const DtwShape& shape, const double* X, int T, int R, int n_pairs,
                         const int* active, const void* dirs, size_t stride, double* v, double* wx) {
    BE_DTW_DISPATCH(launch_dtw_backtrack_w, ctx, X, T, R, n_pairs, active, dirs, stride, v, wx);
}

// table fill + path walk of every (problem, realisation) pair of one DBA iteration
int dtw_paths(be_ctx* ctx, const DtwShape& shape, int tie, const double* A, const double* X, int T, int R, int n_pairs,
              const int* active, void* dirs, size_t stride, double* sq, double* v, double* wx) {
    int rc = launch_dtw_dp(ctx, shape, tie, true, A, X, T, R, R, n_pairs, active, dirs, stride, sq);
    if (rc != BE_OK) return rc;
    return launch_dtw_backtrack(ctx, shape, X, T, R, n_pairs, active, dirs, stride, v, wx);
}

}  // namespace

extern "C" {

size_t be_dtw_dba_workspace_bytes(int B, int R, int T) {
    if (B <= 0 || R <= 0 || T <= 0) return 0;
    return dtw_core_bytes(B, R, T);
}

int be_dtw_squared(be_ctx* ctx, const double* A, const double* X, int P, int T, double* sqcost) {
    if (!ctx) return -1;
    if (!A) return -2;
    if (!X) return -3;
    if (P <= 0) return -4;
    if (T <= 0) return -5;
    if (!sqcost) return -6;
    DtwShape shape;
    if (!dtw_shape(T, shape)) return BE_ERR_UNSUPPORTED;
    // pairs p: row sequence A[p], column sequence X[p]  (R = 1, x_group = 1)
    return launch_dtw_dp(ctx, shape, DTW_TIE_DTWA, false, A, X, T, 1, 1, P, nullptr, nullptr, 0, sqcost);
}

int be_dtw_barycenter_averaging_subgradient(be_ctx* ctx, const double* X, int B, int R, int T, int max_iter,
                                            double initial_step_size, double final_step_size, double tol,
                                            const double* init_barycenter, double* barycenter, int* n_iter,
                                            double* cost, void* workspace, size_t workspace_bytes) {
    NvtxRange nvtx_fn("be_dtw_barycenter_averaging_subgradient");
    if (!ctx) return -1;
    if (!X) return -2;
    if (B <= 0) return -3;
    if (R <= 0) return -4;
    if (T <= 0) return -5;
    if (max_iter < 0) return -6;
    if (!barycenter) return -11;
    if (!workspace) return -14;
    DtwShape shape;
    if (!dtw_shape(T, shape)) return BE_ERR_UNSUPPORTED;
    if (workspace_bytes < dtw_core_bytes(B, R, T)) return BE_ERR_WORKSPACE;
    Carver cv(workspace, workspace_bytes);
    DtwBuffers w;
    if (!carve_dtw(cv, B, R, T, shape, w)) return BE_ERR_WORKSPACE;
    int* iters = n_iter ? n_iter : w.n_iter_own;
    double* cost_last = cost ? cost : w.cost_last_own;
    const size_t stride = dtw_dirs_stride(T, shape);
    const int pairs = B * R;
    int rc;
    if (init_barycenter) {
        BE_CUDA(cudaMemcpyAsync(barycenter, init_barycenter, sizeof(double) * (size_t)B * T, cudaMemcpyDeviceToDevice,
                                ctx->stream));
    } else {
        k_dba_init_mean<<<grid1d((size_t)B * T, 256), 256, 0, ctx->stream>>>(X, B, R, T, barycenter);
        BE_LAUNCHED();
    }
    k_dba_state_init<<<grid1d((size_t)B, 256), 256, 0, ctx->stream>>>(B, w.active, w.cost_prev, cost_last, iters,
                                                                     w.n_active);
    BE_LAUNCHED();
    double eta = initial_step_size;
    for (int it = 0; it < max_iter; ++it) {
        if ((rc = dtw_paths(ctx, shape, DTW_TIE_TSLEARN, barycenter, X, T, R, pairs, w.active, w.dirs, stride, w.sq, w.v,
                            w.wx)) != BE_OK)
            return rc;
        {
            Prof p(ctx, F_DBA_UPDATE, 4.0 * R * T * (double)B, (16.0 * R + 16.0) * T * (double)B);
            k_dba_subgradient_update<<<B, 256, 0, ctx->stream>>>(barycenter, w.v, w.wx, w.sq, T, R,
                                                                 2.0 * eta / (double)R, tol, w.active, w.cost_prev,
                                                                 cost_last, iters, w.n_active);
            BE_LAUNCHED();
        }
        eta -= (initial_step_size - final_step_size) / (double)max_iter;
        // every problem done?  Polled once per four iterations: an iteration over finished problems is three
        // launches whose CTAs exit at once, cheaper than draining the launch queue every time
        if ((it & 3) == 3) {
            int n_active = 0;
            BE_CUDA(cudaMemcpyAsync(&n_active, w.n_active, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            BE_CUDA(cudaStreamSynchronize(ctx->stream));
            if (n_active <= 0) break;
        }
    }
    return BE_OK;
}

int be_perform_dba(be_ctx* ctx, const double* X, int B, int R, int T, int n_iterations, double* center, int* medoid,
                   void* workspace, size_t workspace_bytes) {
    NvtxRange nvtx_fn("be_perform_dba");
    if (!ctx) return -1;
    if (!X) return -2;
    if (B <= 0) return -3;
    if (R <= 0 || R > 50) return -4;  // more than 50 series: the reference samples candidates unseeded (dtwa.py:26)
    if (T <= 0) return -5;
    if (n_iterations < 0) return -6;
    if (!center) return -7;
    if (!workspace) return -9;
    DtwShape shape;
    if (!dtw_shape(T, shape)) return BE_ERR_UNSUPPORTED;
    if (workspace_bytes < dtw_core_bytes(B, R, T)) return BE_ERR_WORKSPACE;
    Carver cv(workspace, workspace_bytes);
    DtwBuffers w;
    if (!carve_dtw(cv, B, R, T, shape, w)) return BE_ERR_WORKSPACE;
    const size_t stride = dtw_dirs_stride(T, shape);
    const int pairs = B * R;
    int rc;
    // medoid: all R x R squared DTW distances per problem (pair p = (b, c, k): rows X[b, c], columns X[b, k])
    if ((rc = launch_dtw_dp(ctx, shape, DTW_TIE_DTWA, false, X, X, T, R, R * R, pairs * R, nullptr, nullptr, 0,
                            w.sq)) != BE_OK)
        return rc;
    k_dba_medoid<<<B, 256, 0, ctx->stream>>>(X, w.sq, T, R, center, medoid);
    BE_LAUNCHED();
    for (int it = 0; it < n_iterations; ++it) {
        if ((rc = dtw_paths(ctx, shape, DTW_TIE_DTWA, center, X, T, R, pairs, nullptr, w.dirs, stride, w.sq, w.v,
                            w.wx)) != BE_OK)
            return rc;
        Prof p(ctx, F_DBA_UPDATE, 2.0 * R * T * (double)B, (16.0 * R + 8.0) * T * (double)B);
        k_dba_mean_update<<<grid1d((size_t)B * T, 256), 256, 0, ctx->stream>>>(center, w.v, w.wx, B, T, R);
        BE_LAUNCHED();
    }
    return BE_OK;
}

}  // extern "C"
