// C ABI, second part (included by be_api.cu): a7 sqrtm, a8 Gaussian W2 distance, a9 full-covariance
// barycentre.  Host orchestration over the tile kernels; see sqrtm_kernels.cuh for the algorithm.
#pragma once

namespace {

struct SqrtmBuffers {
    double *Y, *Z, *WY, *WZ, *VY, *VZ;  // [B][Tp][Tp]
    double *Dinv, *Pbuf;
    double *hldY, *hldZ, *mu, *delta;  // [B]
    double *partialY, *partialZ;       // [B][ctas][2]
};

inline size_t sqrtm_ctas(int T) {
    size_t nblk = num_blocks(pad_dim(T));
    return nblk * (nblk + 1);  // lower tiles x 2 half-tiles
}

size_t sqrtm_core_bytes(int B, int T) {
    size_t mat = align_up(padded_matrix_doubles(B, T) * 8, 256);
    return 6 * mat + align_up(dinv_doubles(B, T) * 8, 256) + align_up(pbuf_doubles(B, T) * 8, 256) +
           4 * align_up((size_t)B * 8, 256) + 2 * align_up((size_t)B * sqrtm_ctas(T) * 2 * 8, 256);
}

bool carve_sqrtm(Carver& cv, int B, int T, SqrtmBuffers& w) {
    const size_t nm = padded_matrix_doubles(B, T);
    w.Y = cv.take<double>(nm); w.Z = cv.take<double>(nm); w.WY = cv.take<double>(nm);
    w.WZ = cv.take<double>(nm); w.VY = cv.take<double>(nm); w.VZ = cv.take<double>(nm);
    w.Dinv = cv.take<double>(dinv_doubles(B, T));
    w.Pbuf = cv.take<double>(pbuf_doubles(B, T));
    w.hldY = cv.take<double>(B); w.hldZ = cv.take<double>(B); w.mu = cv.take<double>(B); w.delta = cv.take<double>(B);
    w.partialY = cv.take<double>((size_t)B * sqrtm_ctas(T) * 2);
    w.partialZ = cv.take<double>((size_t)B * sqrtm_ctas(T) * 2);
    return w.partialZ != nullptr;
}

constexpr double DB_SCALING_OFF = 1e-2;  // determinant scaling is dropped once |dY|/|Y| falls below this

// On entry w.Y holds A (padded, both triangles) and w.WY a copy (its lower triangle is read).
// On return w.Y = A^1/2, w.Z = A^-1/2 (both triangles).  Synchronises the stream once per
// iteration to read the convergence measure.  *iters_host = iterations performed.
int sqrtm_padded(be_ctx* ctx, const SqrtmBuffers& w, int B, int T, double tol, int max_iters, int* iters_host,
                 int* info) {
    const int Tp = pad_dim(T), ld = Tp;
    const int ctas = (int)sqrtm_ctas(T);
    const unsigned fill_grid = ctx->sm_count * 8;
    std::vector<double> delta_h((size_t)B, INFINITY);
    int rc;
    k_set_scaled_identity<<<fill_grid, 256, 0, ctx->stream>>>(w.Z, ld, Tp, T, B, 1.0);
    BE_LAUNCHED();
    k_set_scaled_identity<<<fill_grid, 256, 0, ctx->stream>>>(w.VZ, ld, Tp, T, B, 1.0);  // Z0^-1 = I = V V^T
    BE_LAUNCHED();
    BE_CUDA(cudaMemsetAsync(w.hldZ, 0, sizeof(double) * B, ctx->stream));
    BE_CUDA(cudaMemcpyAsync(w.delta, delta_h.data(), sizeof(double) * B, cudaMemcpyHostToDevice, ctx->stream));
    int it = 0;
    for (; it < max_iters;) {
        // Y^-1 = VY VY^T
        if ((rc = potrf_padded(ctx, w.WY, Tp, T, B, w.Dinv, w.Pbuf, w.VY, info)) != BE_OK) return rc;
        k_diag_reduce<1><<<B, 256, 0, ctx->stream>>>(w.WY, ld, Tp, T, w.hldY, 0);
        BE_LAUNCHED();
        if ((rc = trtri_padded(ctx, w.VY, w.WY, Tp, T, B, w.Dinv, w.Pbuf)) != BE_OK) return rc;
        if (it > 0) {  // Z^-1 = VZ VZ^T  (Z0 = I needs no factorisation)
            if ((rc = potrf_padded(ctx, w.WZ, Tp, T, B, w.Dinv, w.Pbuf, w.VZ, info)) != BE_OK) return rc;
            k_diag_reduce<1><<<B, 256, 0, ctx->stream>>>(w.WZ, ld, Tp, T, w.hldZ, 0);
            BE_LAUNCHED();
            if ((rc = trtri_padded(ctx, w.VZ, w.WZ, Tp, T, B, w.Dinv, w.Pbuf)) != BE_OK) return rc;
        }
        k_db_mu<<<grid1d(B, 128), 128, 0, ctx->stream>>>(w.hldY, w.hldZ, w.delta, DB_SCALING_OFF, T, B, w.mu);
        BE_LAUNCHED();
        // half-tile CTAs past the matrix edge exit without writing their partial sums
        BE_CUDA(cudaMemsetAsync(w.partialY, 0, sizeof(double) * 2 * (size_t)B * ctas, ctx->stream));
        BE_CUDA(cudaMemsetAsync(w.partialZ, 0, sizeof(double) * 2 * (size_t)B * ctas, ctx->stream));
        {
            EpiDB e;
            e.cur = w.Y; e.work = w.WY; e.mu = w.mu; e.partial = w.partialY; e.ld = ld; e.Tp = Tp; e.T = T;
            e.ctas_per_problem = ctas; e.d2 = 0.0; e.n2 = 0.0;
            if ((rc = launch_gemm(ctx, gemm_args(w.VZ, w.VZ, Tp, B, SHAPE_LOWER, KLO_TA, KHI_END, T), e)) != BE_OK) return rc;
        }
        {
            EpiDB e;
            e.cur = w.Z; e.work = w.WZ; e.mu = w.mu; e.partial = w.partialZ; e.ld = ld; e.Tp = Tp; e.T = T;
            e.ctas_per_problem = ctas; e.d2 = 0.0; e.n2 = 0.0;
            if ((rc = launch_gemm(ctx, gemm_args(w.VY, w.VY, Tp, B, SHAPE_LOWER, KLO_TA, KHI_END, T), e)) != BE_OK) return rc;
        }
        k_db_delta<<<grid1d(B, 128), 128, 0, ctx->stream>>>(w.partialY, ctas, B, w.delta);
        BE_LAUNCHED();
        BE_CUDA(cudaMemcpyAsync(delta_h.data(), w.delta, sizeof(double) * B, cudaMemcpyDeviceToHost, ctx->stream));
        BE_CUDA(cudaStreamSynchronize(ctx->stream));
        ++it;
        // A problem whose delta is NaN / inf is not SPD (its info says where): it is left out of the convergence
        // measure and stays NaN, while the other problems of the batch keep iterating until THEY converge.
        double worst = 0.0;
        int good = 0;
        for (int b = 0; b < B; ++b) {
            if (!(delta_h[b] == delta_h[b]) || delta_h[b] == INFINITY) continue;
            ++good;
            if (delta_h[b] > worst) worst = delta_h[b];
        }
        if (good == 0 || worst < tol) break;
    }
    // problems that end above the tolerance (max_iters reached, or NaN) and carry no Cholesky report get their own code
    k_mark_unconverged<<<grid1d(B, 128), 128, 0, ctx->stream>>>(w.delta, tol, B, info);
    BE_LAUNCHED();
    if (iters_host) *iters_host = it;
    return BE_OK;
}

inline GemmArgs gemm_args_div(const double* A, int divA, const double* Bm, int divB, int Tp, int B, int shape, int T = -1) {
    GemmArgs g = gemm_args(A, Bm, Tp, B, shape, KLO_ZERO, KHI_END, T);
    g.divA = divA;
    g.divB = divB;
    return g;
}

}  // namespace

extern "C" {

size_t be_sqrtm_psd_workspace_bytes(int B, int T) { return sqrtm_core_bytes(B, T) + 4096; }

int be_sqrtm_psd(be_ctx* ctx, const double* A, int B, int T, double tol, int max_iters, double* sqrt_out,
                 double* inv_sqrt_out, int* iters_host, int* info, void* workspace, size_t workspace_bytes) {
    NvtxRange nvtx_fn("be_sqrtm_psd");
    if (!ctx) return -1;
    if (!A) return -2;
    if (B <= 0) return -3;
    if (T <= 0) return -4;
    if (!(tol > 0.0)) return -5;
    if (max_iters <= 0) return -6;
    if (!sqrt_out) return -7;
    if (!info) return -10;
    if (!workspace || workspace_bytes < be_sqrtm_psd_workspace_bytes(B, T)) return BE_ERR_WORKSPACE;
    const int Tp = pad_dim(T);
    Carver cv(workspace, workspace_bytes);
    SqrtmBuffers w;
    if (!carve_sqrtm(cv, B, T, w)) return BE_ERR_WORKSPACE;
    BE_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * B, ctx->stream));
    k_pad_full<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(A, B, T, Tp, Tp, w.Y, w.WY);
    BE_LAUNCHED();
    int rc = sqrtm_padded(ctx, w, B, T, tol, max_iters, iters_host, info);
    if (rc != BE_OK) return rc;
    k_copy_out_full<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(w.Y, Tp, Tp, T, sqrt_out, B);
    BE_LAUNCHED();
    if (inv_sqrt_out) {
        k_copy_out_full<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(w.Z, Tp, Tp, T, inv_sqrt_out, B);
        BE_LAUNCHED();
    }
    return BE_OK;
}

size_t be_w2_distance_workspace_bytes(int P, int T) {
    size_t mat = align_up(padded_matrix_doubles(P, T) * 8, 256);
    return sqrtm_core_bytes(P, T) + 2 * mat + 3 * align_up((size_t)P * 8, 256) + 4096;
}

int be_w2_distance(be_ctx* ctx, const double* mu1, const double* sigma1, const double* mu2, const double* sigma2, int P,
                   int T, double tol, int max_iters, double* w2, int* info, void* workspace, size_t workspace_bytes) {
    NvtxRange nvtx_fn("be_w2_distance");
    if (!ctx) return -1;
    if (!mu1) return -2;
    if (!sigma1) return -3;
    if (!mu2) return -4;
    if (!sigma2) return -5;
    if (P <= 0) return -6;
    if (T <= 0) return -7;
    if (!(tol > 0.0)) return -8;
    if (max_iters <= 0) return -9;
    if (!w2) return -10;
    if (!info) return -11;
    if (!workspace || workspace_bytes < be_w2_distance_workspace_bytes(P, T)) return BE_ERR_WORKSPACE;
    const int Tp = pad_dim(T), ld = Tp;
    const size_t nm = padded_matrix_doubles(P, T);
    Carver cv(workspace, workspace_bytes);
    SqrtmBuffers w;
    if (!carve_sqrtm(cv, P, T, w)) return BE_ERR_WORKSPACE;
    double* S2 = cv.take<double>(nm);
    double* G = cv.take<double>(nm);
    double* tr1 = cv.take<double>(P);
    double* tr2 = cv.take<double>(P);
    double* trq = cv.take<double>(P);
    if (!trq) return BE_ERR_WORKSPACE;
    const unsigned fill_grid = ctx->sm_count * 8;
    int rc;
    BE_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * P, ctx->stream));
    // sigma1_sqrt = sqrtm(sigma1)                                         wasserstein.py:41
    k_pad_full<<<fill_grid, 256, 0, ctx->stream>>>(sigma1, P, T, Tp, ld, w.Y, w.WY);
    BE_LAUNCHED();
    k_diag_reduce<0><<<P, 256, 0, ctx->stream>>>(w.Y, ld, Tp, T, tr1, 0);
    BE_LAUNCHED();
    if ((rc = sqrtm_padded(ctx, w, P, T, tol, max_iters, nullptr, info)) != BE_OK) return rc;
    k_pad_full<<<fill_grid, 256, 0, ctx->stream>>>(sigma2, P, T, Tp, ld, S2, nullptr);
    BE_LAUNCHED();
    k_diag_reduce<0><<<P, 256, 0, ctx->stream>>>(S2, ld, Tp, T, tr2, 0);
    BE_LAUNCHED();
    // sigma1_sqrt @ sigma2 @ sigma1_sqrt                                  wasserstein.py:43
    {
        EpiPlain e;
        e.out = G; e.ld = ld; e.Tp = Tp; e.T = T;
        if ((rc = launch_gemm(ctx, gemm_args_div(w.Y, 1, S2, 1, Tp, P, SHAPE_FULL, T), e)) != BE_OK) return rc;  // G = R1 S2
    }
    {
        EpiSym e;
        e.out = S2; e.work = w.WY; e.ld = ld; e.Tp = Tp; e.T = T;
        if ((rc = launch_gemm(ctx, gemm_args_div(w.Y, 1, G, 1, Tp, P, SHAPE_LOWER, T), e)) != BE_OK) return rc;  // R1 S2 R1
    }
    SqrtmBuffers w2b = w;
    w2b.Y = S2;
    if ((rc = sqrtm_padded(ctx, w2b, P, T, tol, max_iters, nullptr, info)) != BE_OK) return rc;
    k_diag_reduce<0><<<P, 256, 0, ctx->stream>>>(S2, ld, Tp, T, trq, 0);
    BE_LAUNCHED();
    k_w2_finish<<<P, 256, 0, ctx->stream>>>(mu1, mu2, T, tr1, tr2, trq, w2);  // wasserstein.py:40,45
    BE_LAUNCHED();
    return BE_OK;
}

int be_w2_distance_diag(be_ctx* ctx, const double* mu1, const double* var1, const double* mu2, const double* var2, int P,
                        int T, double* w2) {
    if (!ctx) return -1;
    if (!mu1) return -2;
    if (!var1) return -3;
    if (!mu2) return -4;
    if (!var2) return -5;
    if (P <= 0) return -6;
    if (T <= 0) return -7;
    if (!w2) return -8;
    k_w2_diag<<<P, 256, 0, ctx->stream>>>(mu1, var1, mu2, var2, T, w2);
    BE_LAUNCHED();
    return BE_OK;
}

size_t be_barycentre_fullcov_workspace_bytes(int C, int M, int T) {
    size_t matB = align_up(padded_matrix_doubles(C * M, T) * 8, 256);
    size_t matC = align_up(padded_matrix_doubles(C, T) * 8, 256);
    return sqrtm_core_bytes(C * M, T) + 2 * matB + 2 * matC + 2 * align_up((size_t)C * 8, 256) +
           align_up((size_t)C * 4, 256) + 4096;
}

int be_barycentre_fullcov(be_ctx* ctx, const double* mus, const double* sigmas, const double* weights, int C, int M,
                          int T, double tolerance, double init_var, int max_iters, double sqrtm_tol,
                          int sqrtm_max_iters, double* mu, double* S_out, int* iters_host, int* info, void* workspace,
                          size_t workspace_bytes) {
    NvtxRange nvtx_fn("be_barycentre_fullcov");
    if (!ctx) return -1;
    if (!mus) return -2;
    if (!sigmas) return -3;
    if (!weights) return -4;
    if (C <= 0) return -5;
    if (M <= 0) return -6;
    if (T <= 0) return -7;
    if (!(init_var > 0.0)) return -9;
    if (max_iters < 0) return -10;
    if (!(sqrtm_tol > 0.0)) return -11;
    if (sqrtm_max_iters <= 0) return -12;
    if (!mu) return -13;
    if (!S_out) return -14;
    if (!info) return -16;
    if (!workspace || workspace_bytes < be_barycentre_fullcov_workspace_bytes(C, M, T)) return BE_ERR_WORKSPACE;
    const int Tp = pad_dim(T), ld = Tp, B = C * M;
    const size_t per = (size_t)Tp * Tp;
    Carver cv(workspace, workspace_bytes);
    SqrtmBuffers w;
    if (!carve_sqrtm(cv, B, T, w)) return BE_ERR_WORKSPACE;
    double* Sig = cv.take<double>(per * B);  // padded members
    double* G = cv.take<double>(per * B);
    double* S = cv.take<double>(per * C);   // current barycentre covariance
    double* Sh = cv.take<double>(per * C);  // its square root
    double* tr_old = cv.take<double>(C);
    double* tr_new = cv.take<double>(C);
    int* active_d = cv.take<int>(C);
    if (!active_d) return BE_ERR_WORKSPACE;
    const unsigned fill_grid = ctx->sm_count * 8;
    int rc;
    BE_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * B, ctx->stream));
    k_weighted_mean<<<grid1d((size_t)C * T, 256), 256, 0, ctx->stream>>>(mus, weights, C, M, T, mu);  // wasserstein.py:98
    BE_LAUNCHED();
    k_pad_full<<<fill_grid, 256, 0, ctx->stream>>>(sigmas, B, T, Tp, ld, Sig, nullptr);
    BE_LAUNCHED();
    k_set_scaled_identity<<<fill_grid, 256, 0, ctx->stream>>>(S, ld, Tp, T, C, init_var);  // wasserstein.py:82
    BE_LAUNCHED();
    std::vector<int> active((size_t)C, 1), n_iters((size_t)C, 0);
    std::vector<double> t_old((size_t)C), t_new((size_t)C);
    int n_active = C;
    while (n_active > 0) {
        BE_CUDA(cudaMemcpyAsync(active_d, active.data(), sizeof(int) * C, cudaMemcpyHostToDevice, ctx->stream));
        // S^1/2
        BE_CUDA(cudaMemcpyAsync(w.Y, S, sizeof(double) * per * C, cudaMemcpyDeviceToDevice, ctx->stream));
        BE_CUDA(cudaMemcpyAsync(w.WY, S, sizeof(double) * per * C, cudaMemcpyDeviceToDevice, ctx->stream));
        if ((rc = sqrtm_padded(ctx, w, C, T, sqrtm_tol, sqrtm_max_iters, nullptr, info)) != BE_OK) return rc;
        BE_CUDA(cudaMemcpyAsync(Sh, w.Y, sizeof(double) * per * C, cudaMemcpyDeviceToDevice, ctx->stream));
        // (S^1/2 Sigma_m S^1/2)^1/2 for every member
        {
            EpiPlain e;
            e.out = G; e.ld = ld; e.Tp = Tp; e.T = T;
            if ((rc = launch_gemm(ctx, gemm_args_div(Sh, M, Sig, 1, Tp, B, SHAPE_FULL, T), e)) != BE_OK) return rc;
        }
        {
            EpiSym e;
            e.out = w.Y; e.work = w.WY; e.ld = ld; e.Tp = Tp; e.T = T;
            if ((rc = launch_gemm(ctx, gemm_args_div(Sh, M, G, 1, Tp, B, SHAPE_LOWER, T), e)) != BE_OK) return rc;
        }
        if ((rc = sqrtm_padded(ctx, w, B, T, sqrtm_tol, sqrtm_max_iters, nullptr, info)) != BE_OK) return rc;
        // candidate = sum_m w_m (.)^1/2 ; signed stop rule on tr(candidate - S) / T       wasserstein.py:85-92
        k_diag_reduce<0><<<C, 256, 0, ctx->stream>>>(S, ld, Tp, T, tr_old, 0);
        BE_LAUNCHED();
        k_weighted_sum<<<fill_grid, 256, 0, ctx->stream>>>(w.Y, weights, active_d, C, M, Tp, ld, T, S, nullptr);
        BE_LAUNCHED();
        k_diag_reduce<0><<<C, 256, 0, ctx->stream>>>(S, ld, Tp, T, tr_new, 0);
        BE_LAUNCHED();
        BE_CUDA(cudaMemcpyAsync(t_old.data(), tr_old, sizeof(double) * C, cudaMemcpyDeviceToHost, ctx->stream));
        BE_CUDA(cudaMemcpyAsync(t_new.data(), tr_new, sizeof(double) * C, cudaMemcpyDeviceToHost, ctx->stream));
        BE_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int c = 0; c < C; ++c) {
            if (!active[c]) continue;
            if ((t_new[c] - t_old[c]) / (double)T < tolerance) {
                active[c] = 0;
            } else {
                n_iters[c] += 1;
                if (n_iters[c] > max_iters) active[c] = 0;  // "not converged": the reference only warns (:94-97)
            }
            if (!active[c]) --n_active;
        }
    }
    if (iters_host)
        for (int c = 0; c < C; ++c) iters_host[c] = n_iters[c];
    k_copy_out_full<<<fill_grid, 256, 0, ctx->stream>>>(S, ld, Tp, T, S_out, C);
    BE_LAUNCHED();
    return BE_OK;
}

/* ---- SURVEY 8f "next": CRPSWeight and ModelSimilarityWeight ------------------------------------ */
int be_crps_weights(be_ctx* ctx, const double* loc, const double* scale, const double* obs, int C, int M, int Ro,
                    int N, double* weights, double* crps_mean) {
    NvtxRange nvtx_fn("be_crps_weights");
    if (!ctx) return -1;
    if (!loc) return -2;
    if (!scale) return -3;
    if (!obs) return -4;
    if (C <= 0) return -5;
    if (M <= 0) return -6;
    if (Ro <= 0) return -7;
    if (N <= 0) return -8;
    if (!weights) return -9;
    const int wb = 128;
    const size_t osm = in_stage_bytes(Ro, wb);
    const size_t tab_bytes = (size_t)2 * (CRPS_TAB_N + 1) * sizeof(double);  // the (G, E) table ahead of the observations
    k_crps_weights<<<grid1d((size_t)C * N, wb), wb, tab_bytes + osm, ctx->stream>>>(loc, scale, obs, C, M, Ro, N, weights,
                                                                                   crps_mean, osm > 0);
    BE_LAUNCHED();
    return BE_OK;
}

int be_ksd_weights(be_ctx* ctx, const double* loc, const double* scale, const double* obs, int C, int M, int Ro,
                   int N, double* weights, double* ksd) {
    NvtxRange nvtx_fn("be_ksd_weights");
    if (!ctx) return -1;
    if (!loc) return -2;
    if (!scale) return -3;
    if (!obs) return -4;
    if (C <= 0) return -5;
    if (M <= 0) return -6;
    if (Ro <= 0) return -7;
    if (N <= 0) return -8;
    if (!weights) return -9;
    const int wb = 128;
    const size_t osm = in_stage_bytes(Ro, wb);
    k_ksd_weights<<<grid1d((size_t)C * N, wb), wb, osm, ctx->stream>>>(loc, scale, obs, C, M, Ro, N, weights, ksd,
                                                                      osm > 0);
    BE_LAUNCHED();
    return BE_OK;
}

int be_w2_collapse(be_ctx* ctx, const double* w2, int C, int M, int N, double* weights) {
    if (!ctx) return -1;
    if (!w2) return -2;
    if (C <= 0) return -3;
    if (M <= 0) return -4;
    if (N <= 0) return -5;
    if (!weights) return -6;
    const int wb = weight_stage_block(M);
    const size_t wsm = weight_stage_bytes(M);
    k_w2_collapse<<<grid1d((size_t)C * N, wb), wb, wsm, ctx->stream>>>(w2, C, M, N, weights, wsm > 0);
    BE_LAUNCHED();
    return BE_OK;
}

int be_similarity_weights_pointwise(be_ctx* ctx, const double* mean, const double* var, int C, int M, int N,
                                    double* weights, double* w2_out) {
    NvtxRange nvtx_fn("be_similarity_weights_pointwise");
    if (!ctx) return -1;
    if (!mean) return -2;
    if (!var) return -3;
    if (C <= 0) return -4;
    if (M <= 0) return -5;
    if (N <= 0) return -6;
    if (!weights) return -7;
    const int wb = 128;
    const size_t msm = in_stage_bytes(M, wb);
    k_similarity_pointwise<<<grid1d((size_t)C * N, wb), wb, msm, ctx->stream>>>(mean, var, C, M, N, weights, w2_out,
                                                                               msm > 0);
    BE_LAUNCHED();
    return BE_OK;
}

}  // extern "C"
