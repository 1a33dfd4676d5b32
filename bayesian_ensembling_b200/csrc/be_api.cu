// C ABI of the B200 hot path (see include/be_b200.h).  Host-side orchestration only: every
// arithmetic operation is a kernel from be_kernels.cuh.  There is no CPU fallback.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/be_b200.h"
#include "be_kernels.cuh"
#include "small_posterior.cuh"
#include "vgp_kernels.cuh"
#include "sqrtm_kernels.cuh"
#include "weights_next_kernels.cuh"
#include "dtw_kernels.cuh"
#include "svgp_kernels.cuh"

using namespace be;

// kernel families of the profiler (be_ctx_profile_*): one per kernel of be_kernels.cuh
enum Family {
    F_INPUTS = 0, F_GRAM, F_DIAG, F_PANEL, F_SYRK, F_TRTRI, F_LAUUM, F_MEAN, F_STATS, F_COPY, F_WEIGHTS, F_BARY,
    F_GEMM, F_VGP_MISC, F_SQRTM_MISC, F_DTW_DP, F_DTW_BACK, F_DBA_UPDATE, F_SMALL_A, F_SMALL_B, F_COUNT
};
static const char* const kFamilyName[F_COUNT] = {
    "k_gpdtw1d_inputs", "k_matern32", "k_diag_block", "k_panel_scale", "k_chol_update", "k_trtri_accum",
    "k_lauum_cov", "k_posterior_mean", "k_mvn_stats", "copy/pad", "k_loglik_weights", "k_barycentre",
    "k_gemm_nt", "vgp elementwise", "sqrtm elementwise", "k_dtw_dp", "k_dtw_backtrack", "k_dba_update",
    "k_small_factor_inverse", "k_small_cov_factor"};

struct ProfRecord {
    int family;
    cudaEvent_t e0, e1;
    double flops, bytes;
};

struct be_ctx {
    int device;
    cudaStream_t stream;
    int sm_count;
    long long launches;
    char err[256];
    bool profiling;
    std::vector<ProfRecord> records;   // launches bracketed since the last reset
    std::vector<cudaEvent_t> ev_pool;  // recycled events
    double fam_ms[F_COUNT], fam_flops[F_COUNT], fam_bytes[F_COUNT];
    long long fam_launches[F_COUNT];
};

namespace {

inline int cuda_fail(be_ctx* ctx, cudaError_t e, const char* what) {
    snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, cudaGetErrorString(e));
    return BE_ERR_CUDA;
}

#define BE_CUDA(call)                                              \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

#define BE_LAUNCHED()                                                   \
    do {                                                                \
        ctx->launches++;                                                \
        cudaError_t e__ = cudaGetLastError();                           \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, "launch");   \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Carver {
    char* base;
    size_t off, cap;
    Carver(void* p, size_t bytes) : base((char*)p), off(0), cap(bytes) {}
    template <typename T>
    T* take(size_t n) {
        size_t bytes = align_up(n * sizeof(T), 256);
        if (off + bytes > cap) return nullptr;
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
};

// Brackets ONE kernel launch with CUDA events on the ctx stream when profiling is on, and books
// the launch's ALGORITHMIC flops / HBM bytes (DESIGN.md "Roofline accounting") to its family.
struct Prof {
    be_ctx* ctx;
    ProfRecord rec;
    bool on;
    Prof(be_ctx* c, int family, double flops, double bytes) : ctx(c), on(c->profiling) {
        if (!on) return;
        rec.family = family;
        rec.flops = flops;
        rec.bytes = bytes;
        rec.e0 = take();
        rec.e1 = take();
        cudaEventRecord(rec.e0, ctx->stream);
    }
    ~Prof() {
        if (!on) return;
        cudaEventRecord(rec.e1, ctx->stream);
        ctx->records.push_back(rec);
    }
    cudaEvent_t take() {
        cudaEvent_t e;
        if (!ctx->ev_pool.empty()) {
            e = ctx->ev_pool.back();
            ctx->ev_pool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        return e;
    }
};

// NVTX range per stage of the path (tracing hook); header-only nvtx3, a no-op without a profiler attached
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// rows / columns of block kb that are REAL (index < T): flops are booked on T, not on the padded Tp
inline double real_width(int T, int kb) {
    int w = T - kb * NB;
    return (double)(w < 0 ? 0 : (w > NB ? NB : w));
}

inline unsigned grid1d(size_t n, int block) { return (unsigned)((n + block - 1) / block); }

// cudaFuncSetAttribute applies to the CURRENT device: once per device, not once per process (a process that
// opens a ctx on a second device must opt its kernels in there too)
bool g_attr_done[64] = {};
int ensure_kernel_attrs(be_ctx* ctx) {
    const int slot = ctx->device >= 0 && ctx->device < 64 ? ctx->device : 0;
    if (g_attr_done[slot]) return BE_OK;
    BE_CUDA(cudaFuncSetAttribute(k_chol_update, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_panel_scale, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_trtri_accum, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_lauum_cov, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_small_factor_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_small_cov_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_diag_block, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_diag_block2, cudaFuncAttributeMaxDynamicSharedMemorySize, DG2_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_matern32<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BE_CUDA(cudaFuncSetAttribute(k_matern32<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BE_CUDA(cudaFuncSetAttribute(k_matern32<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BE_CUDA(cudaFuncSetAttribute(k_gemm_nt<EpiStore>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_gemm_nt<EpiNatP>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_gemm_nt<EpiPhiGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_kbar_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BE_CUDA(cudaFuncSetAttribute(k_gemm_nt<EpiCov>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_gemm_nt<EpiDB>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_gemm_nt<EpiSym>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_gemm_nt<EpiPlain>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_loglik_weights_mvn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WEIGHT_STAGE_MAX_BYTES + 16384));
    BE_CUDA(cudaFuncSetAttribute(k_loglik_weights_mvn_tab<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WEIGHT_STAGE_MAX_BYTES + 16384 + 128));
    BE_CUDA(cudaFuncSetAttribute(k_loglik_weights_mvn_tab<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WEIGHT_STAGE_MAX_BYTES + 16384 + 128));
    BE_CUDA(cudaFuncSetAttribute(k_loglik_weights_normal, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WEIGHT_STAGE_MAX_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_mvn_logprob_vectors, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BE_CUDA(cudaFuncSetAttribute(k_crps_weights, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WEIGHT_STAGE_MAX_BYTES + 16384));
    BE_CUDA(cudaFuncSetAttribute(k_ksd_weights, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WEIGHT_STAGE_MAX_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_w2_collapse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WEIGHT_STAGE_MAX_BYTES));
    BE_CUDA(cudaFuncSetAttribute(k_similarity_pointwise, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WEIGHT_STAGE_MAX_BYTES));
    g_attr_done[slot] = true;
    return BE_OK;
}

inline size_t matern_smem(int R) { return ((size_t)2 * NB * R + 2 * NB) * sizeof(double); }

// Blocked right-looking Cholesky of B padded matrices, in place (lower).  Never pivots on
// columns >= T; rows >= T ride along (file header of be_kernels.cuh).  Fills Dinv with the
// inverted diagonal blocks and, if V != nullptr, the diagonal tiles of V = C^-T.
int potrf_padded(be_ctx* ctx, double* Mat, int Tp, int T, int B, double* Dinv, double* Pbuf, double* V, int* info) {
    const int ld = Tp, nblk = num_blocks(Tp);
    // left-looking: column update (one long-K tensor-core GEMM per tile) -> diagonal block -> panel
    NvtxRange nvtx("be:potrf_padded");
    for (int kb = 0; kb < nblk; ++kb) {
        // algorithmic sizes on the REAL dimension T (the padding and the two right-hand-side rows are not booked)
        const double kw = real_width(T, kb);
        const double nrem = (double)T - kb * NB - kw > 0 ? (double)T - kb * NB - kw : 0.0;  // rows below the diagonal block
        const double kdone = kw > 0 ? (double)kb * NB : 0.0;                              // columns already factorised
        int t = nblk - kb - 1;
        {
            // algorithmic: (nrem x kw) gemm + (kw x kw) syrk, K = kdone (kb == 0: moves the panel to Pbuf)
            Prof pr(ctx, F_SYRK, B * (2.0 * nrem * kw + kw * (kw + 1.0)) * kdone,
                    B * ((nrem + kw) * kdone + 2.0 * (nrem + 0.5 * kw) * kw) * 8);
            k_chol_update<<<(unsigned)((size_t)(t + 1) * 2 * B), GEMM_THREADS, GEMM_SMEM_BYTES, ctx->stream>>>(
                Mat, Pbuf, ld, Tp, kb, B);
            BE_LAUNCHED();
        }
        {
            Prof pr(ctx, F_DIAG, B * (2.0 / 3.0) * kw * kw * kw, B * 3.0 * kw * kw * 8);
            // two forms of the same kernel: with fewer problems than would fill every SM twice, one CTA per SM
            // and the faster single-CTA form (170 KB of shared memory); beyond that the 100 KB form, two CTAs per
            // SM covering each other's latencies (cfg4: 7.36 -> 5.78 ms per step; cfg2, 144 problems: 2.74 vs 3.22)
            if (B < 2 * ctx->sm_count)
                k_diag_block<<<B, 256, DIAG_SMEM_BYTES, ctx->stream>>>(Mat, ld, Tp, T, kb, Dinv, nblk, V, info);
            else
                k_diag_block2<<<B, 256, DG2_SMEM_BYTES, ctx->stream>>>(Mat, ld, Tp, T, kb, Dinv, nblk, V, info);
            BE_LAUNCHED();
        }
        if (t > 0) {
            Prof pr(ctx, F_PANEL, B * nrem * kw * kw, B * (2.0 * nrem * kw + kw * kw) * 8);
            k_panel_scale<<<(unsigned)((size_t)t * 2 * B), GEMM_THREADS, GEMM_SMEM_BYTES, ctx->stream>>>(
                Pbuf, Mat, ld, Tp, kb + 1, kb, Dinv, nblk, 1.0, B);
            BE_LAUNCHED();
        }
    }
    return BE_OK;
}

// V = C^-T (upper, row-major); diagonal tiles already written by potrf_padded.
int trtri_padded(be_ctx* ctx, double* V, const double* Cm, int Tp, int T, int B, const double* Dinv, double* Pbuf) {
    const int ld = Tp, nblk = num_blocks(Tp);
    NvtxRange nvtx("be:trtri_padded");
    for (int i = 1; i < nblk; ++i) {
        const double kw = real_width(T, i);                // booked on T, not Tp
        const double above = kw > 0 ? (double)i * NB : 0;  // rows of V above block i
        {
            // algorithmic: triangular (above x above, upper) times (above x kw): above^2 * kw flops
            Prof pr(ctx, F_TRTRI, B * above * above * kw, B * (0.5 * above * above + 2.0 * above * kw) * 8);
            k_trtri_accum<<<(unsigned)((size_t)i * 2 * B), GEMM_THREADS, GEMM_SMEM_BYTES, ctx->stream>>>(V, Cm, Pbuf, ld,
                                                                                                       Tp, i, B);
            BE_LAUNCHED();
        }
        {
            Prof pr(ctx, F_PANEL, B * above * kw * kw, B * (2.0 * above * kw + kw * kw) * 8);
            k_panel_scale<<<(unsigned)((size_t)i * 2 * B), GEMM_THREADS, GEMM_SMEM_BYTES, ctx->stream>>>(
                Pbuf, V, ld, Tp, 0, i, Dinv, nblk, -1.0, B);
            BE_LAUNCHED();
        }
    }
    return BE_OK;
}

size_t padded_matrix_doubles(int B, int T) {
    size_t Tp = pad_dim(T);
    return (size_t)B * Tp * Tp;
}
size_t dinv_doubles(int B, int T) { return (size_t)B * num_blocks(pad_dim(T)) * NB * NB; }
size_t pbuf_doubles(int B, int T) { return (size_t)B * pad_dim(T) * NB; }

template <class Epi>
int launch_gemm(be_ctx* ctx, const GemmArgs& g, const Epi& epi) {
    const unsigned grid = (unsigned)((size_t)gemm_tiles(g.nblk, g.shape) * 2 * g.B);
    // flops of the contraction ranges at block granularity (structural zeros skipped per 128-block), clipped to the
    // real dimension T: the padding is not booked
    double macs = 0.0;
    for (int tA = 0; tA < g.nblk; ++tA)
        for (int tB = 0; tB < g.nblk; ++tB) {
            if ((g.shape == SHAPE_LOWER && tA < tB) || (g.shape == SHAPE_UPPER && tA > tB)) continue;
            int k0 = g.klo == KLO_ZERO ? 0 : (g.klo == KLO_TA ? tA : (tA > tB ? tA : tB)) * NB;
            int k1 = g.khi == KHI_END ? g.Tp : ((g.khi == KHI_TB ? tB : tA) + 1) * NB;
            if (k1 > g.T) k1 = g.T;
            if (k1 > k0) macs += real_width(g.T, tA) * real_width(g.T, tB) * (double)(k1 - k0);
        }
    Prof pr(ctx, F_GEMM, 2.0 * g.B * macs, 0.0);
    k_gemm_nt<Epi><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, ctx->stream>>>(g, epi);
    BE_LAUNCHED();
    return BE_OK;
}

inline GemmArgs gemm_args(const double* A, const double* Bm, int Tp, int B, int shape, int klo, int khi, int T = -1) {
    GemmArgs g;
    g.A = A;
    g.Bm = Bm;
    g.lda = g.ldb = Tp;
    g.strideA = g.strideB = (size_t)Tp * Tp;
    g.divA = g.divB = 1;
    g.Tp = Tp;
    g.T = T > 0 ? T : Tp;  // real dimension, for the flop booking only
    g.nblk = num_blocks(Tp);
    g.B = B;
    g.shape = shape;
    g.klo = klo;
    g.khi = khi;
    return g;
}

// The L2 loop's padded dimension: at the small-T sizes it is the small-T kernels' (a multiple of 32), so that the two
// factor-and-invert steps of an iteration can run in k_small_factor_inverse (small_posterior.cuh); every other kernel
// of the loop takes the padded dimension as an argument.
inline bool small_path(int T);
inline int vgp_pad(int T) { return small_path(T) ? small_dim(T) : pad_dim(T); }

struct VgpBuffers {
    double *Mk, *Ut, *Wt, *P, *M2, *VP, *VL, *S, *Zt;  // [B][Tp][Tp] each
    double *DinvL, *DinvP, *Pbuf;
    double *n1, *qmu, *r, *v, *zeros;  // [B][T]
    double *u, *am, *av, *partial;
    int *step, *info_tmp;
};

// One natural-gradient step (models.py:209) followed, if train, by one Adam step (models.py:210).
int vgp_iteration(be_ctx* ctx, const VgpBuffers& w, const double* X, const double* y_mean, const double* y_var,
                  double* variance, double* lengthscale, double jitter, double gamma, double lr, int train, int B, int T,
                  int R, int* info_fit) {
    const int Tp = vgp_pad(T), ld = Tp, nblk = num_blocks(Tp);
    const bool small = small_path(T);
    const int ntl = nblk * (nblk + 1) / 2;
    const int nt32 = (Tp + 31) / 32;
    const unsigned rows_grid = grid1d((size_t)B * T, 8);
    int rc;
    // L = chol(K + jitter I)
    k_matern32<1><<<(unsigned)((size_t)ntl * B), 256, matern_smem(R), ctx->stream>>>(X, B, T, R, variance, lengthscale,
                                                                                   w.zeros, w.zeros, jitter, w.Mk, Tp,
                                                                                   ld, ntl);
    BE_LAUNCHED();
    if (small) {
        // L = chol(K + jitter I) AND VL = L^-T in one per-problem kernel (its right-hand-side row is zero here and its
        // posterior-mean tail writes scratch); the lower 32-blocks of VL inside the 128-tiles were zeroed once, before
        // the loop, and nothing writes them
        k_small_factor_inverse<<<B, SM_THREADS, SM_SMEM_BYTES, ctx->stream>>>(w.Mk, w.VL, w.r, info_fit, Tp, T, w.zeros,
                                                                            w.zeros, 0.0, w.v);
        BE_LAUNCHED();
    } else if ((rc = potrf_padded(ctx, w.Mk, Tp, T, B, w.DinvL, w.Pbuf, w.VL, info_fit)) != BE_OK) {
        return rc;
    }
    // Ut = L^T, Wt = L^T D^-1
    k_transpose<<<(unsigned)((size_t)nt32 * nt32 * B), 256, 0, ctx->stream>>>(w.Mk, ld, Tp, T, 1, w.Ut, w.Wt, y_var, B);
    BE_LAUNCHED();
    // theta_1 <- (1-gamma) theta_1 + gamma L^T D^-1 y
    k_rowdot<<<rows_grid, 256, 0, ctx->stream>>>(w.Ut, ld, Tp, T, 2, 1, nullptr, y_mean, y_var, gamma, w.n1, B);
    BE_LAUNCHED();
    // P <- (1-gamma) P + gamma (I + L^T D^-1 L)
    {
        EpiNatP e;
        e.P = w.P; e.work = w.M2; e.G = train ? w.Zt : nullptr; e.ld = ld; e.Tp = Tp; e.T = T; e.gamma = gamma;
        if ((rc = launch_gemm(ctx, gemm_args(w.Wt, w.Ut, Tp, B, SHAPE_LOWER, KLO_TA, KHI_END, T), e)) != BE_OK) return rc;
    }
    // S = P^-1 (potrf, trtri, lauum), q_mu = S theta_1
    if (small) {
        k_small_factor_inverse<<<B, SM_THREADS, SM_SMEM_BYTES, ctx->stream>>>(w.M2, w.VP, w.r, w.info_tmp, Tp, T, w.zeros,
                                                                            w.zeros, 0.0, w.v);
        BE_LAUNCHED();
    } else {
        if ((rc = potrf_padded(ctx, w.M2, Tp, T, B, w.DinvP, w.Pbuf, w.VP, w.info_tmp)) != BE_OK) return rc;
        if ((rc = trtri_padded(ctx, w.VP, w.M2, Tp, T, B, w.DinvP, w.Pbuf)) != BE_OK) return rc;
    }
    {
        EpiStore e;
        e.out = w.S; e.sub = nullptr; e.ld = ld; e.Tp = Tp; e.T = T; e.pad_diag = 1.0; e.mirror = 1;
        if ((rc = launch_gemm(ctx, gemm_args(w.VP, w.VP, Tp, B, SHAPE_LOWER, KLO_TA, KHI_END, T), e)) != BE_OK) return rc;
    }
    k_rowdot<<<rows_grid, 256, 0, ctx->stream>>>(w.S, ld, Tp, T, 0, 0, w.n1, nullptr, nullptr, 0.0, w.qmu, B);
    BE_LAUNCHED();
    if (!train) return BE_OK;
    // r = D^-1 (y - L q_mu)
    k_rowdot<<<rows_grid, 256, 0, ctx->stream>>>(w.Mk, ld, Tp, T, 1, 2, w.qmu, y_mean, y_var, 0.0, w.r, B);
    BE_LAUNCHED();
    // v = L^T r;  Phi = tril(v q_mu^T - G S), halved diagonal, G = L^T D^-1 L kept by the natural-gradient half in Zt
    // (M2 is free again)
    k_rowdot<<<rows_grid, 256, 0, ctx->stream>>>(w.Ut, ld, Tp, T, 2, 0, w.r, nullptr, nullptr, 0.0, w.v, B);
    BE_LAUNCHED();
    {
        EpiPhiGS e;
        e.out = w.M2; e.v = w.v; e.q_mu = w.qmu; e.ld = ld; e.Tp = Tp; e.T = T;
        if ((rc = launch_gemm(ctx, gemm_args(w.Zt, w.S, Tp, B, SHAPE_LOWER, KLO_ZERO, KHI_END, T), e)) != BE_OK) return rc;
    }
    // VL = L^-T  (the small-T kernel formed it with L)
    if (!small && (rc = trtri_padded(ctx, w.VL, w.Mk, Tp, T, B, w.DinvL, w.Pbuf)) != BE_OK) return rc;
    // YT = VL Phi^T  (block upper), into Wt
    {
        EpiStore e;
        e.out = w.Wt; e.sub = nullptr; e.ld = ld; e.Tp = Tp; e.T = T; e.pad_diag = 0.0; e.mirror = 0;
        if ((rc = launch_gemm(ctx, gemm_args(w.VL, w.M2, Tp, B, SHAPE_UPPER, KLO_TA, KHI_TB, T), e)) != BE_OK) return rc;
    }
    // g = sum Kbar_u .* dK/dtheta with Kbar_u = VL YT^T  (into Zt: G is no longer needed), reduced by k_kbar_grad
    const int ctas = nblk * nblk;
    {
        EpiStore e;
        e.out = w.Zt; e.sub = nullptr; e.ld = ld; e.Tp = Tp; e.T = T; e.pad_diag = 0.0; e.mirror = 0;
        if ((rc = launch_gemm(ctx, gemm_args(w.VL, w.Wt, Tp, B, SHAPE_FULL, KLO_MAX, KHI_END, T), e)) != BE_OK) return rc;
    }
    k_kbar_grad<<<(unsigned)((size_t)ctas * B), 256, matern_smem(R), ctx->stream>>>(w.Zt, ld, Tp, X, B, T, R, variance,
                                                                                  lengthscale, w.partial, nblk);
    BE_LAUNCHED();
    k_vgp_adam<<<grid1d(B, 128), 128, 0, ctx->stream>>>(w.partial, ctas, B, lr, 0.9, 0.999, 1e-7, w.u, w.am, w.av, w.step,
                                                       variance, lengthscale);
    BE_LAUNCHED();
    return BE_OK;
}


// ---- small-T member path (small_posterior.cuh): T + 2 <= 256 ---------------------------------------------------
// BE_NO_SMALL_T (environment) sends every size through the blocked path: the A/B switch of the parity tests.
inline bool small_path(int T) {
    const bool off = getenv("BE_NO_SMALL_T") != nullptr;  // read per call: the tests flip it inside one process
    return !off && small_dim(T) <= SM_MAX_DIM;
}
inline size_t small_matrix_doubles(int B, int T) { return (size_t)B * small_dim(T) * small_dim(T); }

size_t gp_posterior_small_workspace_bytes(int B, int T) {
    return 2 * align_up(small_matrix_doubles(B, T) * 8, 256) + align_up((size_t)B * T * 8, 256) + 1024;
}

int gp_posterior_small(be_ctx* ctx, const double* X, const double* y_mean, const double* y_var, const double* variance,
                       const double* lengthscale, double jitter, int B, int T, int R, double* mu, double* var_diag,
                       double* cov, double* scale_tri, double* mvn_stats, int* info_fit, int* info_dist,
                       void* workspace, size_t workspace_bytes) {
    const int n = small_dim(T);
    Carver cv(workspace, workspace_bytes);
    double* Mw = cv.take<double>(small_matrix_doubles(B, T));  // M -> C -> cov -> scale_tri
    double* Vw = cv.take<double>(small_matrix_doubles(B, T));  // V = C^-T
    double* u = cv.take<double>((size_t)B * T);
    if (!Mw || !Vw || !u) return BE_ERR_WORKSPACE;
    BE_CUDA(cudaMemsetAsync(info_fit, 0, sizeof(int) * B, ctx->stream));
    BE_CUDA(cudaMemsetAsync(info_dist, 0, sizeof(int) * B, ctx->stream));
    const int nblk = num_blocks(n), ntl = nblk * (nblk + 1) / 2;
    const double dT = (double)T, dB = (double)B;
    {
        NvtxRange nv("be:small:gram");
        Prof pr(ctx, F_GRAM, dB * 0.5 * dT * dT * (2.0 * R + 12.0), (dB * dT * R + dB * 0.5 * dT * dT) * 8);
        k_matern32<1><<<(unsigned)((size_t)ntl * B), 256, matern_smem(R), ctx->stream>>>(
            X, B, T, R, variance, lengthscale, y_mean, y_var, jitter, Mw, n, n, ntl);
        BE_LAUNCHED();
    }
    {
        // potrf (T^3/3) + triangular inverse (T^3/3); one pass over M and V in global memory (L2-resident)
        NvtxRange nv("be:small:factor_inverse");
        Prof pr(ctx, F_SMALL_A, dB * 2.0 / 3.0 * dT * dT * dT, dB * 1.5 * dT * dT * 8);
        // ... and the posterior mean mu = y - E V u in the kernel's tail (V is L2-hot there)
        k_small_factor_inverse<<<B, SM_THREADS, SM_SMEM_BYTES, ctx->stream>>>(Mw, Vw, u, info_fit, n, T, y_mean, y_var,
                                                                            jitter, mu);
        BE_LAUNCHED();
    }
    {
        // lauum (T^3/3) + Cholesky of the covariance (T^3/3)
        NvtxRange nv("be:small:cov_factor");
        Prof pr(ctx, F_SMALL_B, dB * 2.0 / 3.0 * dT * dT * dT, dB * dT * dT * (cov ? 2.5 : 1.5) * 8);
        k_small_cov_factor<<<B, SM_THREADS, SM_SMEM_BYTES, ctx->stream>>>(Vw, Mw, y_var, jitter, mu, var_diag, cov,
                                                                        info_dist, n, T);
        BE_LAUNCHED();
    }
    {
        Prof pr(ctx, F_STATS, 8.0 * dB * dT, dB * 3.0 * dT * 8);
        k_mvn_stats<<<B, 256, 0, ctx->stream>>>(Mw, n, n, T, mvn_stats);
        BE_LAUNCHED();
    }
    if (scale_tri) {
        Prof pr(ctx, F_COPY, 0.0, dB * 1.5 * dT * dT * 8);
        k_copy_out_tri<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(Mw, n, n, T, scale_tri, B);
        BE_LAUNCHED();
    }
    return BE_OK;
}

}  // namespace

extern "C" {

int be_version(void) { return 100; }

#ifdef BE_SMALL_TIMING
/* developer build only (tools/gpu_small_timing.sh): per-phase cycle totals of one CTA of the small-T kernels */
int be_debug_small_timing(long long* out, int reset) {
    if (out) cudaMemcpyFromSymbol(out, g_small_timing, sizeof(long long) * 2 * 256);
    if (reset) {
        static long long zeros[2 * 256];
        cudaMemcpyToSymbol(g_small_timing, zeros, sizeof(zeros));
    }
    return 0;
}
#endif

int be_ctx_create(int device, void* stream, be_ctx** out) {
    if (!out) return -3;
    be_ctx* ctx = new (std::nothrow) be_ctx();
    if (!ctx) return BE_ERR_CUDA;
    ctx->device = device;
    ctx->stream = (cudaStream_t)stream;
    ctx->launches = 0;
    ctx->err[0] = 0;
    ctx->profiling = false;
    for (int f = 0; f < F_COUNT; ++f) ctx->fam_ms[f] = ctx->fam_flops[f] = ctx->fam_bytes[f] = 0.0, ctx->fam_launches[f] = 0;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        delete ctx;
        return BE_ERR_CUDA;
    }
    int rc = ensure_kernel_attrs(ctx);
    if (rc != BE_OK) {
        fprintf(stderr, "be_ctx_create: %s\n", ctx->err);
        delete ctx;
        return rc;
    }
    *out = ctx;
    return BE_OK;
}

int be_ctx_set_stream(be_ctx* ctx, void* stream) {
    if (!ctx) return -1;
    ctx->stream = (cudaStream_t)stream;
    return BE_OK;
}

int be_ctx_destroy(be_ctx* ctx) {
    if (ctx) {
        for (auto& r : ctx->records) {
            cudaEventDestroy(r.e0);
            cudaEventDestroy(r.e1);
        }
        for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    }
    delete ctx;
    return BE_OK;
}

int be_ctx_profile_enable(be_ctx* ctx, int on) {
    if (!ctx) return -1;
    ctx->profiling = on != 0;
    return BE_OK;
}

// folds the pending event pairs into the per-family totals (synchronises the stream)
static int profile_collect(be_ctx* ctx) {
    if (ctx->records.empty()) return BE_OK;
    BE_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto& r : ctx->records) {
        float ms = 0.f;
        BE_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
        ctx->fam_ms[r.family] += ms;
        ctx->fam_flops[r.family] += r.flops;
        ctx->fam_bytes[r.family] += r.bytes;
        ctx->fam_launches[r.family] += 1;
        ctx->ev_pool.push_back(r.e0);
        ctx->ev_pool.push_back(r.e1);
    }
    ctx->records.clear();
    return BE_OK;
}

int be_ctx_profile_reset(be_ctx* ctx) {
    if (!ctx) return -1;
    int rc = profile_collect(ctx);
    for (int f = 0; f < F_COUNT; ++f) ctx->fam_ms[f] = ctx->fam_flops[f] = ctx->fam_bytes[f] = 0.0, ctx->fam_launches[f] = 0;
    return rc;
}

int be_ctx_profile_families(void) { return F_COUNT; }

int be_ctx_profile_get(be_ctx* ctx, int family, char* name, size_t name_len, double* ms_total,
                       long long* launches, double* flops, double* bytes) {
    if (!ctx) return -1;
    if (family < 0 || family >= F_COUNT) return -2;
    int rc = profile_collect(ctx);
    if (rc != BE_OK) return rc;
    if (name && name_len) {
        strncpy(name, kFamilyName[family], name_len - 1);
        name[name_len - 1] = 0;
    }
    if (ms_total) *ms_total = ctx->fam_ms[family];
    if (launches) *launches = ctx->fam_launches[family];
    if (flops) *flops = ctx->fam_flops[family];
    if (bytes) *bytes = ctx->fam_bytes[family];
    return BE_OK;
}

int be_ctx_sync(be_ctx* ctx) {
    if (!ctx) return -1;
    BE_CUDA(cudaStreamSynchronize(ctx->stream));
    return BE_OK;
}

const char* be_ctx_last_error(be_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }
long long be_ctx_launch_count(be_ctx* ctx) { return ctx ? ctx->launches : 0; }

int be_gpdtw1d_inputs(be_ctx* ctx, const double* realisations, int B, int R, int T, double* X, double* y_mean,
                      double* y_var) {
    NvtxRange nvtx_fn("be_gpdtw1d_inputs");
    if (!ctx) return -1;
    if (!realisations) return -2;
    if (B <= 0) return -3;
    if (R <= 0) return -4;
    if (T <= 0) return -5;
    Prof pr(ctx, F_INPUTS, 4.0 * B * R * T, (2.0 * B * R * T + 2.0 * B * T) * 8);
    k_gpdtw1d_inputs<<<grid1d((size_t)B * T, 256), 256, 0, ctx->stream>>>(realisations, B, R, T, X, y_mean, y_var);
    BE_LAUNCHED();
    return BE_OK;
}

int be_matern32_gram(be_ctx* ctx, const double* X, int B, int T, int R, const double* variance,
                     const double* lengthscale, double* K) {
    NvtxRange nvtx_fn("be_matern32_gram");
    if (!ctx) return -1;
    if (!X) return -2;
    if (B <= 0) return -3;
    if (T <= 0) return -4;
    if (R <= 0 || matern_smem(R) > 200 * 1024) return -5;
    if (!variance) return -6;
    if (!lengthscale) return -7;
    if (!K) return -8;
    int nt = (T + NB - 1) / NB;
    Prof pr(ctx, F_GRAM, (double)B * T * T * (2.0 * R + 12.0), ((double)B * T * R + (double)B * T * T) * 8);
    k_matern32<0><<<(unsigned)((size_t)nt * nt * B), 256, matern_smem(R), ctx->stream>>>(
        X, B, T, R, variance, lengthscale, nullptr, nullptr, 0.0, K, 0, 0, nt * nt);
    BE_LAUNCHED();
    return BE_OK;
}

size_t be_potrf_workspace_bytes(int B, int T) {
    return align_up(padded_matrix_doubles(B, T) * 8, 256) + align_up(dinv_doubles(B, T) * 8, 256) +
           align_up(pbuf_doubles(B, T) * 8, 256) + 1024;
}

int be_potrf_batched(be_ctx* ctx, const double* A, int B, int T, double* L, int* info, void* workspace,
                     size_t workspace_bytes) {
    NvtxRange nvtx_fn("be_potrf_batched");
    if (!ctx) return -1;
    if (!A) return -2;
    if (B <= 0) return -3;
    if (T <= 0) return -4;
    if (!L) return -5;
    if (!info) return -6;
    if (!workspace || workspace_bytes < be_potrf_workspace_bytes(B, T)) return BE_ERR_WORKSPACE;
    const int Tp = pad_dim(T);
    Carver cv(workspace, workspace_bytes);
    double* W = cv.take<double>(padded_matrix_doubles(B, T));
    double* Dinv = cv.take<double>(dinv_doubles(B, T));
    double* Pbuf = cv.take<double>(pbuf_doubles(B, T));
    if (!W || !Dinv || !Pbuf) return BE_ERR_WORKSPACE;
    BE_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * B, ctx->stream));
    k_pad_from_dense<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(A, nullptr, B, T, Tp, Tp, W, 0);
    BE_LAUNCHED();
    int rc = potrf_padded(ctx, W, Tp, T, B, Dinv, Pbuf, nullptr, info);
    if (rc != BE_OK) return rc;
    k_copy_out_tri<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(W, Tp, Tp, T, L, B);
    BE_LAUNCHED();
    return BE_OK;
}

size_t be_gp_posterior_workspace_bytes(int B, int T, int R) {
    (void)R;
    size_t Tp = pad_dim(T);
    size_t blocked = 2 * align_up(padded_matrix_doubles(B, T) * 8, 256) + align_up(dinv_doubles(B, T) * 8, 256) +
                     align_up(pbuf_doubles(B, T) * 8, 256) + align_up((size_t)B * Tp * 8, 256) + 1024;
    if (small_dim(T) <= SM_MAX_DIM) {  // either path may run (BE_NO_SMALL_T): room for both
        size_t small = gp_posterior_small_workspace_bytes(B, T);
        return small > blocked ? small : blocked;
    }
    return blocked;
}

int be_gp_posterior(be_ctx* ctx, const double* X, const double* y_mean, const double* y_var, const double* variance,
                    const double* lengthscale, double jitter, int B, int T, int R, double* mu, double* var_diag,
                    double* cov, double* scale_tri, double* mvn_stats, int* info_fit, int* info_dist, void* workspace,
                    size_t workspace_bytes) {
    NvtxRange nvtx_fn("be_gp_posterior");
    if (!ctx) return -1;
    if (!X) return -2;
    if (!y_mean) return -3;
    if (!y_var) return -4;
    if (!variance) return -5;
    if (!lengthscale) return -6;
    if (!(jitter >= 0.0)) return -7;
    if (B <= 0) return -8;
    if (T <= 0) return -9;
    if (R <= 0 || matern_smem(R) > 200 * 1024) return -10;
    if (!mu) return -11;
    if (!var_diag) return -12;
    if (!mvn_stats) return -15;
    if (!info_fit) return -16;
    if (!info_dist) return -17;
    if (!workspace || workspace_bytes < be_gp_posterior_workspace_bytes(B, T, R)) return BE_ERR_WORKSPACE;
    if (small_path(T))  // T + 2 <= 256: one CTA per member problem, two fused kernels (small_posterior.cuh)
        return gp_posterior_small(ctx, X, y_mean, y_var, variance, lengthscale, jitter, B, T, R, mu, var_diag, cov,
                                  scale_tri, mvn_stats, info_fit, info_dist, workspace, workspace_bytes);
    const int Tp = pad_dim(T), ld = Tp, nblk = num_blocks(Tp);
    Carver cv(workspace, workspace_bytes);
    double* Mw = cv.take<double>(padded_matrix_doubles(B, T));  // M -> C -> cov -> scale_tri
    double* Vw = cv.take<double>(padded_matrix_doubles(B, T));  // V = C^-T
    double* Dinv = cv.take<double>(dinv_doubles(B, T));
    double* Pbuf = cv.take<double>(pbuf_doubles(B, T));
    double* u = cv.take<double>((size_t)B * Tp);
    if (!Mw || !Vw || !Dinv || !Pbuf || !u) return BE_ERR_WORKSPACE;
    BE_CUDA(cudaMemsetAsync(info_fit, 0, sizeof(int) * B, ctx->stream));
    BE_CUDA(cudaMemsetAsync(info_dist, 0, sizeof(int) * B, ctx->stream));

    // 1. M = K + D + jitter I (lower tiles), row T = y_mean
    const int ntl = nblk * (nblk + 1) / 2;
    const double dT = (double)T, dB = (double)B;
    {
        // lower triangle only: T^2/2 entries computed and written
        Prof pr(ctx, F_GRAM, dB * 0.5 * dT * dT * (2.0 * R + 12.0), (dB * dT * R + dB * 0.5 * dT * dT) * 8);
        k_matern32<1><<<(unsigned)((size_t)ntl * B), 256, matern_smem(R), ctx->stream>>>(
            X, B, T, R, variance, lengthscale, y_mean, y_var, jitter, Mw, Tp, ld, ntl);
        BE_LAUNCHED();
    }
    // 2. C = chol(M); row T becomes u = C^-1 y; V diagonal tiles
    int rc = potrf_padded(ctx, Mw, Tp, T, B, Dinv, Pbuf, Vw, info_fit);
    if (rc != BE_OK) return rc;
    {
        Prof pr(ctx, F_COPY, 0.0, dB * dT * 16);
        k_extract_row<<<grid1d((size_t)B * T, 256), 256, 0, ctx->stream>>>(Mw, ld, Tp, T, T, u, B, 1);
        BE_LAUNCHED();
    }
    // 3. V = C^-T
    rc = trtri_padded(ctx, Vw, Mw, Tp, T, B, Dinv, Pbuf);
    if (rc != BE_OK) return rc;
    // 4. mean = y - E V u
    {
        Prof pr(ctx, F_MEAN, dB * dT * dT, dB * (0.5 * dT * dT + 4.0 * dT) * 8);
        k_posterior_mean<<<grid1d((size_t)B * T, 8), 256, 0, ctx->stream>>>(Vw, ld, Tp, T, u, y_mean, y_var, jitter,
                                                                            mu, B);
        BE_LAUNCHED();
    }
    // 5. cov = D + E - E (V V^T) E  -> Mw (padded, rows T/T+1 = 1, mu), var_diag, dense cov
    {
        // lauum: T^3/3 flops; reads V (upper, T^2/2), writes the padded lower cov (+ dense cov if asked)
        Prof pr(ctx, F_LAUUM, dB * dT * dT * dT / 3.0, dB * dT * dT * (cov ? 2.0 : 1.0) * 8);
        const int raster = getenv("BE_LAUUM_TILE_MAJOR") ? 0 : 1;  // A/B switch (tools/gpu_traffic.sh)
        k_lauum_cov<<<(unsigned)((size_t)ntl * 2 * B), GEMM_THREADS, GEMM_SMEM_BYTES, ctx->stream>>>(
            Vw, ld, Tp, T, y_var, jitter, mu, Mw, var_diag, cov, B, raster);
        BE_LAUNCHED();
    }
    // 6. scale_tri = chol(cov) (data.py:38-39); rows T/T+1 become a = L^-1 1, b = L^-1 mu
    rc = potrf_padded(ctx, Mw, Tp, T, B, Dinv, Pbuf, nullptr, info_dist);
    if (rc != BE_OK) return rc;
    {
        Prof pr(ctx, F_STATS, 8.0 * dB * dT, dB * 3.0 * dT * 8);
        k_mvn_stats<<<B, 256, 0, ctx->stream>>>(Mw, ld, Tp, T, mvn_stats);
        BE_LAUNCHED();
    }
    if (scale_tri) {
        Prof pr(ctx, F_COPY, 0.0, dB * 1.5 * dT * dT * 8);
        k_copy_out_tri<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(Mw, ld, Tp, T, scale_tri, B);
        BE_LAUNCHED();
    }
    return BE_OK;
}

/* Same inputs and (mu, var_diag, mvn_stats, info) outputs as be_gp_posterior without the dense cov / scale_tri:
 * the posterior covariance stays in factored form (Woodbury), T^3 instead of 4/3 T^3 tensor flops. */
size_t be_gp_posterior_factored_workspace_bytes(int B, int T, int R) {
    return be_gp_posterior_workspace_bytes(B, T, R) + 3 * align_up((size_t)B * T * 8, 256) +
           3 * align_up((size_t)B * 4 * 8, 256);
}

int be_gp_posterior_factored(be_ctx* ctx, const double* X, const double* y_mean, const double* y_var,
                             const double* variance, const double* lengthscale, double jitter, int B, int T, int R,
                             double* mu, double* var_diag, double* mvn_stats, int* info_fit, int* info_dist,
                             void* workspace, size_t workspace_bytes) {
    NvtxRange nvtx_fn("be_gp_posterior_factored");
    if (!ctx) return -1;
    if (!X) return -2;
    if (!y_mean) return -3;
    if (!y_var) return -4;
    if (!variance) return -5;
    if (!lengthscale) return -6;
    if (!(jitter >= 0.0)) return -7;
    if (B <= 0) return -8;
    if (T <= 0) return -9;
    if (R <= 0 || matern_smem(R) > 200 * 1024) return -10;
    if (!mu) return -11;
    if (!var_diag) return -12;
    if (!mvn_stats) return -13;
    if (!info_fit) return -14;
    if (!info_dist) return -15;
    if (!workspace || workspace_bytes < be_gp_posterior_factored_workspace_bytes(B, T, R)) return BE_ERR_WORKSPACE;
    const int Tp = pad_dim(T), ld = Tp, nblk = num_blocks(Tp);
    Carver cv(workspace, workspace_bytes);
    double* Mw = cv.take<double>(padded_matrix_doubles(B, T));  // M -> C, then N -> L_N
    double* Vw = cv.take<double>(padded_matrix_doubles(B, T));  // V = C^-T
    double* Dinv = cv.take<double>(dinv_doubles(B, T));
    double* Pbuf = cv.take<double>(pbuf_doubles(B, T));
    double* u = cv.take<double>((size_t)B * Tp);
    double* nvar = cv.take<double>((size_t)B * T);
    double* g = cv.take<double>((size_t)B * T);
    double* gmu = cv.take<double>((size_t)B * T);
    double* base = cv.take<double>((size_t)B * 4);
    double* statsC = cv.take<double>((size_t)B * 4);
    double* statsN = cv.take<double>((size_t)B * 4);
    if (!Mw || !Vw || !Dinv || !Pbuf || !u || !nvar || !g || !gmu || !base || !statsC || !statsN) return BE_ERR_WORKSPACE;
    BE_CUDA(cudaMemsetAsync(info_fit, 0, sizeof(int) * B, ctx->stream));
    BE_CUDA(cudaMemsetAsync(info_dist, 0, sizeof(int) * B, ctx->stream));
    const int ntl = nblk * (nblk + 1) / 2;
    const double dT = (double)T, dB = (double)B;
    // 1-2. M = K + E = C C^T, u = C^-1 y rides along; sum log diag C
    {
        Prof pr(ctx, F_GRAM, dB * 0.5 * dT * dT * (2.0 * R + 12.0), (dB * dT * R + dB * 0.5 * dT * dT) * 8);
        k_matern32<1><<<(unsigned)((size_t)ntl * B), 256, matern_smem(R), ctx->stream>>>(
            X, B, T, R, variance, lengthscale, y_mean, y_var, jitter, Mw, Tp, ld, ntl);
        BE_LAUNCHED();
    }
    int rc = potrf_padded(ctx, Mw, Tp, T, B, Dinv, Pbuf, Vw, info_fit);
    if (rc != BE_OK) return rc;
    {
        Prof pr(ctx, F_COPY, 0.0, dB * dT * 16);
        k_extract_row<<<grid1d((size_t)B * T, 256), 256, 0, ctx->stream>>>(Mw, ld, Tp, T, T, u, B, 1);
        BE_LAUNCHED();
    }
    {
        Prof pr(ctx, F_STATS, 8.0 * dB * dT, dB * 3.0 * dT * 8);
        k_mvn_stats<<<B, 256, 0, ctx->stream>>>(Mw, ld, Tp, T, statsC);
        BE_LAUNCHED();
    }
    // 3-4. V = C^-T; mean = y - E V u and var_diag = D + E - E^2 diag(V V^T) in one pass over V
    rc = trtri_padded(ctx, Vw, Mw, Tp, T, B, Dinv, Pbuf);
    if (rc != BE_OK) return rc;
    {
        Prof pr(ctx, F_MEAN, 2.0 * dB * dT * dT, dB * (0.5 * dT * dT + 5.0 * dT) * 8);
        k_posterior_mean<<<grid1d((size_t)B * T, 8), 256, 0, ctx->stream>>>(Vw, ld, Tp, T, u, y_mean, y_var, jitter,
                                                                            mu, B, var_diag);
        BE_LAUNCHED();
    }
    // 5. N = K + diag(E D / E') = L_N L_N^T with the rows (G 1, G mu) riding along
    {
        Prof pr(ctx, F_STATS, 12.0 * dB * dT, dB * 5.0 * dT * 8);
        k_factored_prepare<<<B, 256, 0, ctx->stream>>>(y_var, mu, jitter, T, nvar, g, gmu, base);
        BE_LAUNCHED();
    }
    {
        Prof pr(ctx, F_GRAM, dB * 0.5 * dT * dT * (2.0 * R + 12.0), (dB * dT * R + dB * 0.5 * dT * dT) * 8);
        k_matern32<1><<<(unsigned)((size_t)ntl * B), 256, matern_smem(R), ctx->stream>>>(
            X, B, T, R, variance, lengthscale, g, nvar, 0.0, Mw, Tp, ld, ntl);
        BE_LAUNCHED();
    }
    {
        Prof pr(ctx, F_COPY, 0.0, dB * dT * 16);
        k_set_row<<<grid1d((size_t)B * T, 256), 256, 0, ctx->stream>>>(Mw, ld, Tp, T, T + 1, gmu, B);
        BE_LAUNCHED();
    }
    rc = potrf_padded(ctx, Mw, Tp, T, B, Dinv, Pbuf, nullptr, info_dist);
    if (rc != BE_OK) return rc;
    {
        Prof pr(ctx, F_STATS, 8.0 * dB * dT, dB * 3.0 * dT * 8);
        k_mvn_stats<<<B, 256, 0, ctx->stream>>>(Mw, ld, Tp, T, statsN);
        BE_LAUNCHED();
        k_factored_finish<<<grid1d((size_t)B * 4, 128), 128, 0, ctx->stream>>>(statsN, base, statsC, B, mvn_stats);
        BE_LAUNCHED();
    }
    return BE_OK;
}

size_t be_mvn_from_cov_workspace_bytes(int B, int T) { return be_potrf_workspace_bytes(B, T); }

int be_mvn_from_cov(be_ctx* ctx, const double* mu, const double* cov, int B, int T, double* scale_tri,
                    double* var_diag, double* mvn_stats, int* info, void* workspace, size_t workspace_bytes) {
    NvtxRange nvtx_fn("be_mvn_from_cov");
    if (!ctx) return -1;
    if (!mu) return -2;
    if (!cov) return -3;
    if (B <= 0) return -4;
    if (T <= 0) return -5;
    if (!mvn_stats) return -8;
    if (!info) return -9;
    if (!workspace || workspace_bytes < be_mvn_from_cov_workspace_bytes(B, T)) return BE_ERR_WORKSPACE;
    const int Tp = pad_dim(T);
    Carver cv(workspace, workspace_bytes);
    double* W = cv.take<double>(padded_matrix_doubles(B, T));
    double* Dinv = cv.take<double>(dinv_doubles(B, T));
    double* Pbuf = cv.take<double>(pbuf_doubles(B, T));
    if (!W || !Dinv || !Pbuf) return BE_ERR_WORKSPACE;
    BE_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * B, ctx->stream));
    k_pad_from_dense<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(cov, mu, B, T, Tp, Tp, W, 1);
    BE_LAUNCHED();
    if (var_diag) {
        k_diag_from_dense<<<grid1d((size_t)B * T, 256), 256, 0, ctx->stream>>>(cov, B, T, var_diag);
        BE_LAUNCHED();
    }
    int rc = potrf_padded(ctx, W, Tp, T, B, Dinv, Pbuf, nullptr, info);
    if (rc != BE_OK) return rc;
    k_mvn_stats<<<B, 256, 0, ctx->stream>>>(W, Tp, Tp, T, mvn_stats);
    BE_LAUNCHED();
    if (scale_tri) {
        k_copy_out_tri<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(W, Tp, Tp, T, scale_tri, B);
        BE_LAUNCHED();
    }
    return BE_OK;
}

int be_loglik_weights_mvn(be_ctx* ctx, const double* mvn_stats, const double* obs, int C, int M, int Ro, int T,
                          double standardisation_constant, double* weights, double* lls_exp, double* lls_mean) {
    NvtxRange nvtx_fn("be_loglik_weights_mvn");
    if (!ctx) return -1;
    if (!mvn_stats) return -2;
    if (!obs) return -3;
    if (C <= 0) return -4;
    if (M <= 0) return -5;
    if (Ro <= 0) return -6;
    if (T <= 0) return -7;
    if (!weights) return -9;
    const int nout = 1 + (lls_exp ? 1 : 0) + (lls_mean ? 1 : 0);
    Prof pr(ctx, F_WEIGHTS, (double)C * M * T * (8.0 * Ro + 24.0),
            ((double)C * Ro * T + (double)C * M * 4 + (double)nout * C * M * T) * 8);
    const int wb = weight_stage_block(M);
    const size_t wsm = weight_stage_bytes(M);
    const size_t ssm = weight_stats_bytes(M) <= 16384 ? weight_stats_bytes(M) : 0;
    // BE_WEIGHTS_LIBEXP: the library-exp kernel for every M (the A/B switch of tools/prof_weights_ab.py)
    static const bool lib_exp = getenv("BE_WEIGHTS_LIBEXP") != nullptr;
    if (wsm > 0 && ssm > 0 && !lib_exp) {
        const size_t sm = wsm + ssm + 16 * sizeof(double);
        if (lls_exp || lls_mean)
            k_loglik_weights_mvn_tab<true><<<grid1d((size_t)C * T, wb), wb, sm, ctx->stream>>>(
                mvn_stats, obs, C, M, Ro, T, standardisation_constant, 1.0 / Ro, weights, lls_exp, lls_mean);
        else
            k_loglik_weights_mvn_tab<false><<<grid1d((size_t)C * T, wb), wb, sm, ctx->stream>>>(
                mvn_stats, obs, C, M, Ro, T, standardisation_constant, 1.0 / Ro, weights, nullptr, nullptr);
    } else {
        k_loglik_weights_mvn<<<grid1d((size_t)C * T, wb), wb, wsm + ssm, ctx->stream>>>(
            mvn_stats, obs, C, M, Ro, T, standardisation_constant, weights, lls_exp, lls_mean,
            (wsm > 0 ? 1 : 0) | (ssm > 0 ? 2 : 0));
    }
    BE_LAUNCHED();
    return BE_OK;
}

int be_mvn_constvec_logprob(be_ctx* ctx, const double* mvn_stats, const double* obs, int C, int M, int Ro, int T,
                            double* ll) {
    if (!ctx) return -1;
    if (!mvn_stats) return -2;
    if (!obs) return -3;
    if (C <= 0) return -4;
    if (M <= 0) return -5;
    if (Ro <= 0) return -6;
    if (T <= 0) return -7;
    if (!ll) return -8;
    k_mvn_constvec_logprob<<<grid1d((size_t)C * M * Ro * T, 256), 256, 0, ctx->stream>>>(mvn_stats, obs, C, M, Ro, T,
                                                                                         ll);
    BE_LAUNCHED();
    return BE_OK;
}

int be_mvn_log_prob(be_ctx* ctx, const double* mu, const double* scale_tri, const double* x, int T, int N,
                    double sum_log_diag, double* ll) {
    if (!ctx) return -1;
    if (!mu) return -2;
    if (!scale_tri) return -3;
    if (!x) return -4;
    if (T <= 0 || (size_t)T * sizeof(double) > 200 * 1024) return -5;
    if (N <= 0) return -6;
    if (!ll) return -8;
    k_mvn_logprob_vectors<<<N, 256, (size_t)T * sizeof(double), ctx->stream>>>(mu, scale_tri, x, T, N, sum_log_diag, ll);
    BE_LAUNCHED();
    return BE_OK;
}

int be_normal_logprob(be_ctx* ctx, const double* loc, const double* scale, const double* x, size_t n, double* ll) {
    if (!ctx) return -1;
    if (!loc) return -2;
    if (!scale) return -3;
    if (!x) return -4;
    if (n == 0) return -5;
    if (!ll) return -6;
    k_normal_logprob<<<grid1d(n, 256), 256, 0, ctx->stream>>>(loc, scale, x, n, ll);
    BE_LAUNCHED();
    return BE_OK;
}

int be_loglik_weights_normal(be_ctx* ctx, const double* loc, const double* scale, const double* obs, int C, int M,
                             int Ro, int N, double standardisation_constant, double* weights, double* lls_exp,
                             double* lls_mean) {
    NvtxRange nvtx_fn("be_loglik_weights_normal");
    if (!ctx) return -1;
    if (!loc) return -2;
    if (!scale) return -3;
    if (!obs) return -4;
    if (C <= 0) return -5;
    if (M <= 0) return -6;
    if (Ro <= 0) return -7;
    if (N <= 0) return -8;
    if (!weights) return -10;
    const int wb = weight_stage_block(M);
    const size_t wsm = weight_stage_bytes(M);
    k_loglik_weights_normal<<<grid1d((size_t)C * N, wb), wb, wsm, ctx->stream>>>(
        loc, scale, obs, C, M, Ro, N, standardisation_constant, weights, lls_exp, lls_mean, wsm > 0);
    BE_LAUNCHED();
    return BE_OK;
}

int be_weights_time_mean(be_ctx* ctx, const double* weights, int C, int M, int T, double* w_bar) {
    NvtxRange nvtx_fn("be_weights_time_mean");
    if (!ctx) return -1;
    if (!weights) return -2;
    if (C <= 0) return -3;
    if (M <= 0) return -4;
    if (T <= 0) return -5;
    if (!w_bar) return -6;
    k_weights_time_mean<<<(unsigned)((size_t)C * M), 256, 0, ctx->stream>>>(weights, C * M, T, w_bar);
    BE_LAUNCHED();
    return BE_OK;
}

int be_weights_normalise(be_ctx* ctx, const double* lls_exp, const double* total, int C, int M, int T, double* weights) {
    if (!ctx) return -1;
    if (!lls_exp) return -2;
    if (!total) return -3;
    if (C <= 0) return -4;
    if (M <= 0) return -5;
    if (T <= 0) return -6;
    if (!weights) return -7;
    k_weights_normalise<<<grid1d((size_t)C * M * T, 256), 256, 0, ctx->stream>>>(lls_exp, total, C, M, T, weights);
    BE_LAUNCHED();
    return BE_OK;
}

int be_barycentre_1d(be_ctx* ctx, const double* means, const double* variances, const double* weights, int C, int M,
                     int N, double tolerance, double init_var, int max_iters, double* mu, double* sigma, int* iters) {
    NvtxRange nvtx_fn("be_barycentre_1d");
    if (!ctx) return -1;
    if (!means) return -2;
    if (!variances) return -3;
    if (!weights) return -4;
    if (C <= 0) return -5;
    if (M <= 0) return -6;
    if (N <= 0) return -7;
    if (!mu) return -11;
    if (!sigma) return -12;
    // 24*M + 16 bytes per (cell, time) point
    Prof pr(ctx, F_BARY, 4.0 * C * M * N, (double)C * N * (24.0 * M + 16.0));
    k_barycentre_1d<<<grid1d((size_t)C * N, 128), 128, 0, ctx->stream>>>(means, variances, weights, C, M, N, tolerance,
                                                                         init_var, max_iters, mu, sigma, iters);
    BE_LAUNCHED();
    return BE_OK;
}

int be_barycentre_1d_partial(be_ctx* ctx, const double* means, const double* variances, const double* lls_exp, int C,
                             int M_local, int N, double* partial) {
    if (!ctx) return -1;
    if (!means) return -2;
    if (!variances) return -3;
    if (!lls_exp) return -4;
    if (C <= 0) return -5;
    if (M_local <= 0) return -6;
    if (N <= 0) return -7;
    if (!partial) return -8;
    k_barycentre_partial<<<grid1d((size_t)C * N, 128), 128, 0, ctx->stream>>>(means, variances, lls_exp, C, M_local, N,
                                                                              partial);
    BE_LAUNCHED();
    return BE_OK;
}

int be_barycentre_1d_finish(be_ctx* ctx, const double* partial, int C, int N, double tolerance, double init_var,
                            int max_iters, double* mu, double* sigma, int* iters) {
    if (!ctx) return -1;
    if (!partial) return -2;
    if (C <= 0) return -3;
    if (N <= 0) return -4;
    if (!mu) return -8;
    if (!sigma) return -9;
    k_barycentre_finish<<<grid1d((size_t)C * N, 128), 128, 0, ctx->stream>>>(partial, C, N, tolerance, init_var,
                                                                             max_iters, mu, sigma, iters);
    BE_LAUNCHED();
    return BE_OK;
}

size_t be_vgp_fit_workspace_bytes(int B, int T, int R) {
    (void)R;
    size_t Tp = vgp_pad(T);
    size_t mat = align_up((size_t)B * Tp * Tp * 8, 256);
    size_t vec = align_up((size_t)B * T * 8, 256);
    size_t ctas = (size_t)num_blocks((int)Tp) * num_blocks((int)Tp) * 2;
    size_t dinv = align_up((size_t)B * num_blocks((int)Tp) * NB * NB * 8, 256), pbuf = align_up((size_t)B * Tp * NB * 8, 256);
    return 9 * mat + 2 * dinv + pbuf + 5 * vec +
           3 * align_up((size_t)B * 2 * 8, 256) + align_up((size_t)B * ctas * 2 * 8, 256) +
           2 * align_up((size_t)B * 4, 256) + 4096;
}

int be_vgp_fit(be_ctx* ctx, const double* X, const double* y_mean, const double* y_var, int B, int T, int R,
               int n_iters, double gamma, double lr, int train_hypers, double jitter, double* variance,
               double* lengthscale, double* mu, double* var_diag, double* cov, double* scale_tri, double* mvn_stats,
               int* info_fit, int* info_dist, void* workspace, size_t workspace_bytes) {
    NvtxRange nvtx_fn("be_vgp_fit");
    if (!ctx) return -1;
    if (!X) return -2;
    if (!y_mean) return -3;
    if (!y_var) return -4;
    if (B <= 0) return -5;
    if (T <= 0) return -6;
    if (R <= 0 || matern_smem(R) > 200 * 1024) return -7;
    if (n_iters < 0) return -8;
    if (!(gamma > 0.0 && gamma <= 1.0)) return -9;
    if (!(jitter >= 0.0)) return -12;
    if (!variance) return -13;
    if (!lengthscale) return -14;
    if (!mu) return -15;
    if (!var_diag) return -16;
    if (!cov) return -17;
    if (!mvn_stats) return -19;
    if (!info_fit) return -20;
    if (!info_dist) return -21;
    if (!workspace || workspace_bytes < be_vgp_fit_workspace_bytes(B, T, R)) return BE_ERR_WORKSPACE;
    const int Tp = vgp_pad(T), ld = Tp, nblk = num_blocks(Tp);
    const size_t nm = (size_t)B * Tp * Tp;
    const size_t n_dinv = (size_t)B * nblk * NB * NB, n_pbuf = (size_t)B * Tp * NB;
    Carver cv(workspace, workspace_bytes);
    VgpBuffers w;
    w.Mk = cv.take<double>(nm); w.Ut = cv.take<double>(nm); w.Wt = cv.take<double>(nm); w.P = cv.take<double>(nm);
    w.M2 = cv.take<double>(nm); w.VP = cv.take<double>(nm); w.VL = cv.take<double>(nm); w.S = cv.take<double>(nm);
    w.Zt = cv.take<double>(nm);
    w.DinvL = cv.take<double>(n_dinv); w.DinvP = cv.take<double>(n_dinv);
    w.Pbuf = cv.take<double>(n_pbuf);
    w.n1 = cv.take<double>((size_t)B * T); w.qmu = cv.take<double>((size_t)B * T); w.r = cv.take<double>((size_t)B * T);
    w.zeros = cv.take<double>((size_t)B * T); w.v = cv.take<double>((size_t)B * T);
    w.u = cv.take<double>((size_t)B * 2); w.am = cv.take<double>((size_t)B * 2); w.av = cv.take<double>((size_t)B * 2);
    w.partial = cv.take<double>((size_t)B * nblk * nblk * 2 * 2);
    w.step = cv.take<int>(B); w.info_tmp = cv.take<int>(B);
    if (!w.info_tmp) return BE_ERR_WORKSPACE;

    cudaStream_t user_stream = ctx->stream;
    BE_CUDA(cudaMemsetAsync(info_fit, 0, sizeof(int) * B, user_stream));
    BE_CUDA(cudaMemsetAsync(info_dist, 0, sizeof(int) * B, user_stream));
    BE_CUDA(cudaMemsetAsync(w.info_tmp, 0, sizeof(int) * B, user_stream));
    BE_CUDA(cudaMemsetAsync(w.n1, 0, sizeof(double) * B * T, user_stream));
    BE_CUDA(cudaMemsetAsync(w.qmu, 0, sizeof(double) * B * T, user_stream));
    BE_CUDA(cudaMemsetAsync(w.zeros, 0, sizeof(double) * B * T, user_stream));
    BE_CUDA(cudaMemsetAsync(w.am, 0, sizeof(double) * B * 2, user_stream));
    BE_CUDA(cudaMemsetAsync(w.av, 0, sizeof(double) * B * 2, user_stream));
    BE_CUDA(cudaMemsetAsync(w.step, 0, sizeof(int) * B, user_stream));
    if (small_path(T)) {
        // k_small_factor_inverse writes the 32-blocks of V on and above the diagonal only; the tile products read whole
        // 128-tiles, so the blocks below it have to be zero (once: nothing else writes them inside the loop)
        BE_CUDA(cudaMemsetAsync(w.VL, 0, sizeof(double) * nm, user_stream));
        BE_CUDA(cudaMemsetAsync(w.VP, 0, sizeof(double) * nm, user_stream));
    }
    k_set_identity<<<ctx->sm_count * 8, 256, 0, user_stream>>>(w.P, ld, Tp, B);  // q_sqrt = I
    BE_LAUNCHED();
    k_set_identity<<<ctx->sm_count * 8, 256, 0, user_stream>>>(w.S, ld, Tp, B);
    BE_LAUNCHED();
    k_vgp_unconstrain<<<grid1d(B, 128), 128, 0, user_stream>>>(variance, lengthscale, B, w.u);
    BE_LAUNCHED();

    // The training loop: one iteration is captured into a CUDA graph on a private stream and replayed
    // n_iters times -- ~60 small launches per iteration would otherwise be launch-latency bound.
    if (n_iters > 0) {
        cudaStream_t cap;
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
        int rc = BE_OK;
        if (ce == cudaSuccess) {
            rc = vgp_iteration(ctx, w, X, y_mean, y_var, variance, lengthscale, jitter, gamma, lr, train_hypers, B, T, R,
                               info_fit);
            ce = cudaStreamEndCapture(cap, &graph);
        }
        const long long per_iter = ctx->launches - launches0;
        if (ce == cudaSuccess && rc == BE_OK) ce = cudaGraphInstantiate(&exec, graph, 0);
        if (ce == cudaSuccess && rc == BE_OK) {
            for (int it = 0; it < n_iters && ce == cudaSuccess; ++it) ce = cudaGraphLaunch(exec, cap);
            ctx->launches = launches0 + per_iter * n_iters;
        }
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        ctx->stream = user_stream;
        ctx->profiling = prof;
        cudaError_t ce2 = cudaEventRecord(join, cap);
        if (ce2 == cudaSuccess) ce2 = cudaStreamWaitEvent(user_stream, join, 0);
        cudaEventDestroy(fork);
        cudaEventDestroy(join);
        // the private stream can be destroyed once its work is ordered before user_stream's next op
        cudaStreamDestroy(cap);
        if (rc != BE_OK) return rc;
        if (ce != cudaSuccess) return cuda_fail(ctx, ce, "vgp graph");
        if (ce2 != cudaSuccess) return cuda_fail(ctx, ce2, "vgp join");
    }

    // predict_f(X, full_cov=True) at the final hyper-parameters (models.py:217) + diag(y_var) (:220)
    const int ntl = nblk * (nblk + 1) / 2;
    const int nt32 = (Tp + 31) / 32;
    int rc;
    k_matern32<1><<<(unsigned)((size_t)ntl * B), 256, matern_smem(R), ctx->stream>>>(X, B, T, R, variance, lengthscale,
                                                                                   w.zeros, w.zeros, jitter, w.Mk, Tp,
                                                                                   ld, ntl);
    BE_LAUNCHED();
    if ((rc = potrf_padded(ctx, w.Mk, Tp, T, B, w.DinvL, w.Pbuf, w.VL, info_fit)) != BE_OK) return rc;
    if ((rc = trtri_padded(ctx, w.VL, w.Mk, Tp, T, B, w.DinvL, w.Pbuf)) != BE_OK) return rc;
    k_matern32<2><<<(unsigned)((size_t)nblk * nblk * B), 256, matern_smem(R), ctx->stream>>>(
        X, B, T, R, variance, lengthscale, nullptr, nullptr, 0.0, w.Ut, Tp, ld, nblk * nblk);  // Ut := K (no jitter)
    BE_LAUNCHED();
    k_transpose<<<(unsigned)((size_t)nt32 * nt32 * B), 256, 0, ctx->stream>>>(w.VL, ld, Tp, T, 2, w.Wt, nullptr, nullptr,
                                                                             B);  // Wt := Lm^-1 (lower)
    BE_LAUNCHED();
    {   // AT = K Lm^-T  (= A^T, A = Lm^-1 K)
        EpiStore e;
        e.out = w.Zt; e.sub = nullptr; e.ld = ld; e.Tp = Tp; e.T = T; e.pad_diag = 0.0; e.mirror = 0;
        if ((rc = launch_gemm(ctx, gemm_args(w.Ut, w.Wt, Tp, B, SHAPE_FULL, KLO_ZERO, KHI_TB, T), e)) != BE_OK) return rc;
    }
    k_rowdot<<<grid1d((size_t)B * T, 8), 256, 0, ctx->stream>>>(w.Zt, ld, Tp, T, 0, 0, w.qmu, nullptr, nullptr, 0.0, mu, B);
    BE_LAUNCHED();
    {   // M2 = AT S - AT
        EpiStore e;
        e.out = w.M2; e.sub = w.Zt; e.ld = ld; e.Tp = Tp; e.T = T; e.pad_diag = 0.0; e.mirror = 0;
        if ((rc = launch_gemm(ctx, gemm_args(w.Zt, w.S, Tp, B, SHAPE_FULL, KLO_ZERO, KHI_END, T), e)) != BE_OK) return rc;
    }
    {   // cov = K + (AT (S - I)) AT^T + D
        EpiCov e;
        e.K = w.Ut; e.y_var = y_var; e.cov = cov; e.var_diag = var_diag; e.ld = ld; e.Tp = Tp; e.T = T;
        if ((rc = launch_gemm(ctx, gemm_args(w.M2, w.Zt, Tp, B, SHAPE_LOWER, KLO_ZERO, KHI_END, T), e)) != BE_OK) return rc;
    }
    // Distribution(mu, cov, MultivariateNormalFullCovariance): data.py:38-39
    k_pad_from_dense<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(cov, mu, B, T, Tp, Tp, w.P, 1);
    BE_LAUNCHED();
    if ((rc = potrf_padded(ctx, w.P, Tp, T, B, w.DinvP, w.Pbuf, nullptr, info_dist)) != BE_OK) return rc;
    k_mvn_stats<<<B, 256, 0, ctx->stream>>>(w.P, ld, Tp, T, mvn_stats);
    BE_LAUNCHED();
    if (scale_tri) {
        k_copy_out_tri<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(w.P, ld, Tp, T, scale_tri, B);
        BE_LAUNCHED();
    }
    return BE_OK;
}

}  // extern "C"

#include "be_w2_api.cuh"
#include "be_dtw_api.cuh"
#include "be_svgp_api.cuh"
