// C ABI, fourth part (included by be_api.cu): the SVGP stage of GPDTW3D.fit (ensembles/models.py:357-411).
// Host orchestration over svgp_kernels.cuh and the blocked factorisation of be_api.cu; oracle/svgp.py is the spec.
#pragma once

namespace {

template <int TA, int TB>
int svgp_gemm(be_ctx* ctx, int m, int n, int k, double alpha, const double* A, int lda, const double* Bm, int ldb, double beta,
              double* C, int ldc) {
    k_dgemm<TA, TB><<<dgemm_grid(m, n), 128, 0, ctx->stream>>>(m, n, k, alpha, A, lda, Bm, ldb, beta, C, ldc);
    BE_LAUNCHED();
    return BE_OK;
}

struct SvgpBuffers {
    // M x M matrices in the padded [Mp, Mp] layout
    double *Lu, *Vu, *P, *Wp, *Vp, *S, *Sq;
    double *DinvA, *DinvB, *Pbuf;
    // plain M x M (ld M)
    double *G, *T1, *T2, *LuD, *SqD;  // LuD / SqD: the factors with a clean (zero) strict upper triangle
    // M x n (ld n)
    double *Kuf, *A, *Aw, *W, *SW, *Abar, *Kufbar;
    // vectors
    double *Xb, *yb, *sb, *fmean, *gm, *gv, *n1, *n1s, *qmu, *u, *g, *gZ, *am, *av;
    int *step, *info_tmp;
};

size_t svgp_workspace_bytes(int D, int M, int nmax) {
    const size_t Mp = pad_dim(M);
    const size_t mat = align_up(Mp * Mp * 8, 256), mm = align_up((size_t)M * M * 8, 256), mn = align_up((size_t)M * nmax * 8, 256);
    const size_t dinv = align_up(dinv_doubles(1, M) * 8, 256), pbuf = align_up(pbuf_doubles(1, M) * 8, 256);
    const size_t vec = align_up((size_t)(nmax > M ? nmax : M) * 8, 256);
    return 7 * mat + 2 * dinv + pbuf + 5 * mm + 7 * mn + align_up((size_t)nmax * D * 8, 256) + 8 * vec +
           4 * align_up((size_t)(8 + (size_t)M * D) * 8, 256) + 1024 + 4096;
}

bool svgp_carve(Carver& cv, int D, int M, int nmax, SvgpBuffers& w) {
    const size_t Mp = pad_dim(M), nm = Mp * Mp;
    const size_t vlen = nmax > M ? nmax : M;
    w.Lu = cv.take<double>(nm); w.Vu = cv.take<double>(nm); w.P = cv.take<double>(nm); w.Wp = cv.take<double>(nm);
    w.Vp = cv.take<double>(nm); w.S = cv.take<double>(nm); w.Sq = cv.take<double>(nm);
    w.DinvA = cv.take<double>(dinv_doubles(1, M)); w.DinvB = cv.take<double>(dinv_doubles(1, M));
    w.Pbuf = cv.take<double>(pbuf_doubles(1, M));
    w.G = cv.take<double>((size_t)M * M); w.T1 = cv.take<double>((size_t)M * M); w.T2 = cv.take<double>((size_t)M * M);
    w.LuD = cv.take<double>((size_t)M * M); w.SqD = cv.take<double>((size_t)M * M);
    w.Kuf = cv.take<double>((size_t)M * nmax); w.A = cv.take<double>((size_t)M * nmax); w.Aw = cv.take<double>((size_t)M * nmax);
    w.W = cv.take<double>((size_t)M * nmax); w.SW = cv.take<double>((size_t)M * nmax); w.Abar = cv.take<double>((size_t)M * nmax);
    w.Kufbar = cv.take<double>((size_t)M * nmax);
    w.Xb = cv.take<double>((size_t)nmax * D);
    w.yb = cv.take<double>(vlen); w.sb = cv.take<double>(vlen); w.fmean = cv.take<double>(vlen); w.gm = cv.take<double>(vlen);
    w.gv = cv.take<double>(vlen); w.n1 = cv.take<double>(vlen); w.n1s = cv.take<double>(vlen); w.qmu = cv.take<double>(vlen);
    const size_t np = 8 + (size_t)M * D;
    w.u = cv.take<double>(np); w.gZ = cv.take<double>(np); w.am = cv.take<double>(np); w.av = cv.take<double>(np);
    w.g = cv.take<double>(8);
    w.step = cv.take<int>(1); w.info_tmp = cv.take<int>(1);
    return w.info_tmp != nullptr;
}

// Lu = chol(Kuu + jitter I) (padded, in w.Lu), Vu = Lu^-T (upper, row-major): Lu^-1 = Vu^T
int svgp_factor_kuu(be_ctx* ctx, const SvgpBuffers& w, const double* Z, const SvgpKernelParams& kp, int M, double jitter, int* info) {
    const int Mp = pad_dim(M);
    int rc;
    k_svgp_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(Z, M, Z, M, kp, jitter, Mp, w.Lu, Mp);
    BE_LAUNCHED();
    // the blocked kernels write only the upper block triangle of V: its lower blocks must read as zero in the products
    BE_CUDA(cudaMemsetAsync(w.Vu, 0, sizeof(double) * (size_t)Mp * Mp, ctx->stream));
    if ((rc = potrf_padded(ctx, w.Lu, Mp, M, 1, w.DinvA, w.Pbuf, w.Vu, info)) != BE_OK) return rc;
    if ((rc = trtri_padded(ctx, w.Vu, w.Lu, Mp, M, 1, w.DinvA, w.Pbuf)) != BE_OK) return rc;
    k_copy_out_tri<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(w.Lu, Mp, Mp, M, w.LuD, 1);  // upper blocks still hold Kuu
    BE_LAUNCHED();
    return BE_OK;
}

// A [M, n] = Lu^-1 K(Z, Xb)
int svgp_conditional_A(be_ctx* ctx, const SvgpBuffers& w, const double* Z, const SvgpKernelParams& kp, int M, int n) {
    const int Mp = pad_dim(M);
    k_svgp_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(Z, M, w.Xb, n, kp, 0.0, 0, w.Kuf, n);
    BE_LAUNCHED();
    return svgp_gemm<1, 0>(ctx, M, n, M, 1.0, w.Vu, Mp, w.Kuf, n, 0.0, w.A, n);
}

// One optimisation step (models.py:389-391): natural-gradient half on minibatch 2 * step, Adam half on 2 * step + 1.
int svgp_iteration(be_ctx* ctx, const SvgpBuffers& w, const double* X, const double* Y, const long long* batch_idx, double* Z,
                   double* variances, double* lengthscales, const SvgpKernelParams& kp, int D, int M, int n, double gamma,
                   double lr, int train_hypers, double jitter, int* info) {
    const int Mp = pad_dim(M);
    int rc;
    // ---- natural-gradient step on minibatch 2 * step (models.py:390)
    k_svgp_gather<<<grid1d((size_t)n * D, 256), 256, 0, ctx->stream>>>(X, Y, batch_idx, w.step, 0, n, D, w.Xb, w.yb,
                                                                     w.sb);
    BE_LAUNCHED();
    if ((rc = svgp_factor_kuu(ctx, w, Z, kp, M, jitter, info)) != BE_OK) return rc;
    if ((rc = svgp_conditional_A(ctx, w, Z, kp, M, n)) != BE_OK) return rc;
    k_svgp_scale_cols<<<grid1d((size_t)M * n, 256), 256, 0, ctx->stream>>>(w.A, w.sb, M, n, w.Aw);
    BE_LAUNCHED();
    k_gemv_n<<<grid1d((size_t)M * 32, 256), 256, 0, ctx->stream>>>(M, n, w.Aw, n, w.yb, w.n1s);  // nat1* = A D^-1 y
    BE_LAUNCHED();
    if ((rc = svgp_gemm<0, 1>(ctx, M, M, n, 1.0, w.Aw, n, w.A, n, 0.0, w.G, M)) != BE_OK) return rc;  // A D^-1 A^T
    k_svgp_natgrad_update<<<grid1d((size_t)Mp * Mp, 256), 256, 0, ctx->stream>>>(w.P, w.G, w.n1, w.n1s, M, Mp, gamma, w.Wp);
    BE_LAUNCHED();
    // S = P^-1 = Vp Vp^T (Vp = chol(P)^-T), q_mu = S nat1, q_sqrt = chol(S)
    BE_CUDA(cudaMemsetAsync(w.Vp, 0, sizeof(double) * (size_t)Mp * Mp, ctx->stream));
    if ((rc = potrf_padded(ctx, w.Wp, Mp, M, 1, w.DinvB, w.Pbuf, w.Vp, w.info_tmp)) != BE_OK) return rc;
    if ((rc = trtri_padded(ctx, w.Vp, w.Wp, Mp, M, 1, w.DinvB, w.Pbuf)) != BE_OK) return rc;
    if ((rc = svgp_gemm<0, 1>(ctx, M, M, M, 1.0, w.Vp, Mp, w.Vp, Mp, 0.0, w.S, Mp)) != BE_OK) return rc;
    k_gemv_n<<<grid1d((size_t)M * 32, 256), 256, 0, ctx->stream>>>(M, M, w.S, Mp, w.n1, w.qmu);
    BE_LAUNCHED();
    k_svgp_pad_copy<<<grid1d((size_t)Mp * Mp, 256), 256, 0, ctx->stream>>>(w.S, Mp, M, Mp, w.Sq);
    BE_LAUNCHED();
    if ((rc = potrf_padded(ctx, w.Sq, Mp, M, 1, w.DinvB, w.Pbuf, nullptr, w.info_tmp)) != BE_OK) return rc;
    k_copy_out_tri<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(w.Sq, Mp, Mp, M, w.SqD, 1);
    BE_LAUNCHED();
    if (train_hypers) {
    // ---- Adam step on minibatch 2 * step + 1 (models.py:391): same kernel parameters, so Lu / Vu stand
    k_svgp_gather<<<grid1d((size_t)n * D, 256), 256, 0, ctx->stream>>>(X, Y, batch_idx, w.step, 1, n, D, w.Xb, w.yb,
                                                                     w.sb);
    BE_LAUNCHED();
    if ((rc = svgp_conditional_A(ctx, w, Z, kp, M, n)) != BE_OK) return rc;
    k_gemv_t<<<grid1d(n, 32), 256, 0, ctx->stream>>>(M, n, w.A, n, w.qmu, w.fmean);  // m = A^T q_mu
    BE_LAUNCHED();
    if ((rc = svgp_gemm<1, 0>(ctx, M, n, M, 1.0, w.SqD, M, w.A, n, 0.0, w.W, n)) != BE_OK) return rc;   // W = Sq^T A
    if ((rc = svgp_gemm<0, 0>(ctx, M, n, M, 1.0, w.SqD, M, w.W, n, 0.0, w.SW, n)) != BE_OK) return rc;  // Sq W
    k_svgp_point_grads<<<grid1d(n, 128), 128, 0, ctx->stream>>>(w.fmean, w.yb, w.sb, n, w.gm, w.gv);
    BE_LAUNCHED();
    k_svgp_abar<<<grid1d((size_t)M * n, 256), 256, 0, ctx->stream>>>(w.A, w.SW, w.qmu, w.gm, w.gv, M, n, w.Abar);
    BE_LAUNCHED();
    if ((rc = svgp_gemm<0, 0>(ctx, M, n, M, 1.0, w.Vu, Mp, w.Abar, n, 0.0, w.Kufbar, n)) != BE_OK) return rc;  // Lu^-T Abar
    if ((rc = svgp_gemm<0, 1>(ctx, M, M, n, 1.0, w.Kufbar, n, w.A, n, 0.0, w.T1, M)) != BE_OK) return rc;      // Kuf_bar A^T
    k_svgp_neg_tril<<<grid1d((size_t)M * M, 256), 256, 0, ctx->stream>>>(w.T1, M);                            // Lu_bar
    BE_LAUNCHED();
    if ((rc = svgp_gemm<1, 0>(ctx, M, M, M, 1.0, w.LuD, M, w.T1, M, 0.0, w.T2, M)) != BE_OK) return rc;  // Lu^T Lu_bar
    k_svgp_phi<<<grid1d((size_t)M * M, 256), 256, 0, ctx->stream>>>(w.T2, M);
    BE_LAUNCHED();
    if ((rc = svgp_gemm<0, 0>(ctx, M, M, M, 1.0, w.Vu, Mp, w.T2, M, 0.0, w.T1, M)) != BE_OK) return rc;  // Lu^-T Phi
    if ((rc = svgp_gemm<0, 1>(ctx, M, M, M, 1.0, w.T1, M, w.Vu, Mp, 0.0, w.G, M)) != BE_OK) return rc;   // ... Lu^-1
    BE_CUDA(cudaMemsetAsync(w.g, 0, sizeof(double) * 8, ctx->stream));
    k_svgp_param_grads<<<M, 128, 0, ctx->stream>>>(Z, w.Xb, w.Kufbar, w.G, w.gv, kp, M, n, w.g, w.gZ);
    BE_LAUNCHED();
    k_svgp_adam<<<grid1d(8 + (size_t)M * D, 128), 128, 0, ctx->stream>>>(w.g, w.gZ, M * D, lr, w.u, Z, w.am, w.av, w.step,
                                                                       variances, lengthscales);
    BE_LAUNCHED();
    }
    k_svgp_step_inc<<<1, 1, 0, ctx->stream>>>(w.step);
    BE_LAUNCHED();
    return BE_OK;
}

}  // namespace

extern "C" {

size_t be_svgp_fit_workspace_bytes(int N, int D, int M, int minibatch_size, int predict_chunk) {
    (void)N;
    if (D < 5 || D > SVGP_MAX_D || M <= 0 || minibatch_size <= 0) return 0;
    const int nmax = predict_chunk > minibatch_size ? predict_chunk : minibatch_size;
    return svgp_workspace_bytes(D, M, nmax);
}

int be_svgp_fit(be_ctx* ctx, const double* X, const double* Y, int N, int D, int M, int minibatch_size,
                const long long* batch_idx, int n_steps, double gamma, double lr, int train_hypers, double jitter,
                int predict_chunk, double* Z, double* variances, double* lengthscales, double* q_mu, double* q_sqrt, double* mu,
                double* var, int* info, void* workspace, size_t workspace_bytes) {
    NvtxRange nvtx_fn("be_svgp_fit");
    if (!ctx) return -1;
    if (!X) return -2;
    if (!Y) return -3;
    if (N <= 0) return -4;
    if (D < 5 || D > SVGP_MAX_D) return -5;
    if (M <= 0) return -6;
    if (minibatch_size <= 0) return -7;
    if (n_steps > 0 && !batch_idx) return -8;
    if (n_steps < 0) return -9;
    if (!(gamma > 0.0 && gamma <= 1.0)) return -10;
    if (!(jitter >= 0.0)) return -13;
    if (predict_chunk <= 0) return -14;
    if (!Z) return -15;
    if (!variances) return -16;
    if (!lengthscales) return -17;
    if (!q_mu) return -18;
    if (!q_sqrt) return -19;
    if (!mu) return -20;
    if (!var) return -21;
    if (!info) return -22;
    if (!workspace || workspace_bytes < be_svgp_fit_workspace_bytes(N, D, M, minibatch_size, predict_chunk)) return BE_ERR_WORKSPACE;
    const int n = minibatch_size, Mp = pad_dim(M);
    const int nmax = predict_chunk > n ? predict_chunk : n;
    Carver cv(workspace, workspace_bytes);
    SvgpBuffers w;
    if (!svgp_carve(cv, D, M, nmax, w)) return BE_ERR_WORKSPACE;
    SvgpKernelParams kp{variances, lengthscales, D};
    const unsigned fill = ctx->sm_count * 4;
    int rc;
    BE_CUDA(cudaMemsetAsync(info, 0, sizeof(int), ctx->stream));
    BE_CUDA(cudaMemsetAsync(w.info_tmp, 0, sizeof(int), ctx->stream));
    BE_CUDA(cudaMemsetAsync(w.n1, 0, sizeof(double) * M, ctx->stream));
    BE_CUDA(cudaMemsetAsync(w.qmu, 0, sizeof(double) * M, ctx->stream));
    BE_CUDA(cudaMemsetAsync(w.am, 0, sizeof(double) * (8 + (size_t)M * D), ctx->stream));
    BE_CUDA(cudaMemsetAsync(w.av, 0, sizeof(double) * (8 + (size_t)M * D), ctx->stream));
    BE_CUDA(cudaMemsetAsync(w.step, 0, sizeof(int), ctx->stream));
    k_set_identity<<<fill, 256, 0, ctx->stream>>>(w.P, Mp, Mp, 1);   // S^-1 = I   (q_sqrt = I)
    BE_LAUNCHED();
    k_set_identity<<<fill, 256, 0, ctx->stream>>>(w.Sq, Mp, Mp, 1);  // q_sqrt = I
    BE_LAUNCHED();
    k_copy_out_tri<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(w.Sq, Mp, Mp, M, w.SqD, 1);
    BE_LAUNCHED();
    k_svgp_unconstrain<<<1, 32, 0, ctx->stream>>>(variances, lengthscales, w.u);
    BE_LAUNCHED();

    // The training loop: one step (~90 small launches) is captured into a CUDA graph on a private stream and replayed
    // n_steps times, as be_vgp_fit does; the minibatch offset comes from the device-side step counter.
    if (n_steps > 0) {
        cudaStream_t user_stream = ctx->stream, cap;
        cudaEvent_t fork, join;
        BE_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
        BE_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        BE_CUDA(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
        BE_CUDA(cudaEventRecord(fork, user_stream));
        BE_CUDA(cudaStreamWaitEvent(cap, fork, 0));
        const bool prof = ctx->profiling;
        const long long launches0 = ctx->launches;
        ctx->profiling = false;
        ctx->stream = cap;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaError_t ce = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
        rc = BE_OK;
        if (ce == cudaSuccess) {
            rc = svgp_iteration(ctx, w, X, Y, batch_idx, Z, variances, lengthscales, kp, D, M, n, gamma, lr, train_hypers, jitter,
                                info);
            ce = cudaStreamEndCapture(cap, &graph);
        }
        const long long per_iter = ctx->launches - launches0;
        if (ce == cudaSuccess && rc == BE_OK) ce = cudaGraphInstantiate(&exec, graph, 0);
        if (ce == cudaSuccess && rc == BE_OK) {
            for (int it = 0; it < n_steps && ce == cudaSuccess; ++it) ce = cudaGraphLaunch(exec, cap);
            ctx->launches = launches0 + per_iter * n_steps;
        }
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        ctx->stream = user_stream;
        ctx->profiling = prof;
        cudaError_t ce2 = cudaEventRecord(join, cap);
        if (ce2 == cudaSuccess) ce2 = cudaStreamWaitEvent(user_stream, join, 0);
        cudaEventDestroy(fork);
        cudaEventDestroy(join);
        cudaStreamDestroy(cap);
        if (rc != BE_OK) return rc;
        if (ce != cudaSuccess) return cuda_fail(ctx, ce, "svgp graph");
        if (ce2 != cudaSuccess) return cuda_fail(ctx, ce2, "svgp join");
    }
    // ---- predict_f(X, full_cov=False) at the final parameters (models.py:408), + Y[:, 1] (models.py:411)
    if ((rc = svgp_factor_kuu(ctx, w, Z, kp, M, jitter, info)) != BE_OK) return rc;
    for (long long c0 = 0; c0 < N; c0 += predict_chunk) {
        const int nc = (int)((N - c0) < predict_chunk ? (N - c0) : predict_chunk);
        k_svgp_gather_rows<<<grid1d((size_t)nc * D, 256), 256, 0, ctx->stream>>>(X, Y, c0, nc, D, w.Xb, w.sb);
        BE_LAUNCHED();
        if ((rc = svgp_conditional_A(ctx, w, Z, kp, M, nc)) != BE_OK) return rc;
        k_gemv_t<<<grid1d(nc, 32), 256, 0, ctx->stream>>>(M, nc, w.A, nc, w.qmu, mu + c0);
        BE_LAUNCHED();
        if ((rc = svgp_gemm<1, 0>(ctx, M, nc, M, 1.0, w.SqD, M, w.A, nc, 0.0, w.W, nc)) != BE_OK) return rc;
        k_svgp_predict_var<<<grid1d(nc, 128), 128, 0, ctx->stream>>>(w.A, w.W, variances, w.sb, M, nc, var + c0);
        BE_LAUNCHED();
    }
    BE_CUDA(cudaMemcpyAsync(q_mu, w.qmu, sizeof(double) * M, cudaMemcpyDeviceToDevice, ctx->stream));
    k_copy_out_tri<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(w.Sq, Mp, Mp, M, q_sqrt, 1);
    BE_LAUNCHED();
    return BE_OK;
}

}  // extern "C"
