/*
 * be_b200.h -- C ABI of the B200-native fit -> weight -> barycentre hot path of
 * mattramos/bayesian_ensembling.
 *
 * The reference is pure Python and has NO FFI; its numerics are calls into
 * GPflow/TensorFlow, distrax/JAX and NumPy made from the files cited below
 * (paths relative to the reference repo).  Every entry point here replaces one
 * of those call sites; INTEGRATION.md shows the ctypes binding a maintainer of
 * the reference would add at each site.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / CUDA types in signatures
 *     (a cudaStream_t travels as void*).
 *   - all arrays are fp64, C-order, contiguous, DEVICE pointers unless the
 *     parameter name ends in _host.
 *   - B enumerates independent (cell, member) problems, b = cell * M + member.
 *   - return value: 0 ok; <0 = -(index of the bad argument, 1-based);
 *     BE_ERR_CUDA / BE_ERR_WORKSPACE for runtime failures.  Numerical
 *     conditions (non-PD matrix) are NOT errors: they are reported per problem
 *     in the LAPACK-style `info` arrays (0 ok, k>0 = leading minor k not PD)
 *     and NaNs propagate exactly as in the reference (jnp.linalg.cholesky
 *     returns NaN silently).
 *   - the caller owns every buffer including workspace (query *_workspace_bytes);
 *     the ctx owns only its stream handle.  Calls on one ctx are stream-ordered
 *     and asynchronous; be_ctx_sync() waits.  One ctx per (process, device).
 */
#ifndef BE_B200_H
#define BE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BE_OK 0
#define BE_ERR_CUDA 1000
#define BE_ERR_WORKSPACE 1001
#define BE_ERR_UNSUPPORTED 1002
/* per-problem info of the iterative routines (be_sqrtm_psd, be_w2_distance, be_barycentre_fullcov): > 0 and
 * below this value = first non-positive-definite leading minor (LAPACK style); this value = the iteration ended
 * above its tolerance (max_iters reached or a NaN iterate) -- the result for that problem is not to be used */
#define BE_INFO_NOT_CONVERGED 0x40000000

#define BE_DEFAULT_JITTER 1e-6 /* gpflow.config.default_jitter() */

typedef struct be_ctx be_ctx;

int be_version(void);
/* stream may be NULL (legacy default stream) or a cudaStream_t owned by the caller */
int be_ctx_create(int device, void* stream, be_ctx** out);
int be_ctx_set_stream(be_ctx* ctx, void* stream);
int be_ctx_destroy(be_ctx* ctx);
int be_ctx_sync(be_ctx* ctx);
/* text of the last CUDA error seen on this ctx ("" if none) */
const char* be_ctx_last_error(be_ctx* ctx);
/* number of kernels this ctx has launched since creation (bench.py's gpu_launches) */
long long be_ctx_launch_count(be_ctx* ctx);

/* ---- per-kernel timing (bench.py's roofline object) --------------------------------------
 * While enabled, every kernel launch is bracketed by CUDA events on the ctx stream and its
 * ALGORITHMIC flops / HBM bytes are booked to the kernel's family (DESIGN.md "Roofline
 * accounting").  be_ctx_profile_get synchronises the stream and returns the totals since the
 * last reset for family 0 <= family < be_ctx_profile_families(). */
int be_ctx_profile_enable(be_ctx* ctx, int on);
int be_ctx_profile_reset(be_ctx* ctx);
int be_ctx_profile_families(void);
int be_ctx_profile_get(be_ctx* ctx, int family, char* name, size_t name_len, double* ms_total,
                       long long* launches, double* flops, double* bytes);

/* ---- a1 inputs: ensembles/models.py:175-182 -------------------------------------------
 * realisations [B,R,T] -> X [B,T,R] (= realisation_set.T), y_mean [B,T] (arithmetic mean
 * over realisations; the reference's DBA mean, models.py:176-178, is supplied by the caller
 * instead when wanted), y_var [B,T] (np.var, ddof=0, models.py:179). */
int be_gpdtw1d_inputs(be_ctx* ctx, const double* realisations, int B, int R, int T,
                      double* X, double* y_mean, double* y_var);

/* ---- a1 kernel matrix: gpf.kernels.Matern32() built at models.py:186 -------------------
 * K [B,T,T] dense symmetric; variance/lengthscale [B]. */
int be_matern32_gram(be_ctx* ctx, const double* X, int B, int T, int R,
                     const double* variance, const double* lengthscale, double* K);

/* ---- a3 Cholesky: jnp.linalg.cholesky inside distrax, ensembles/data.py:38-39 ----------
 * A [B,T,T] dense (lower triangle read) -> L [B,T,T] lower, strict upper zeroed. */
size_t be_potrf_workspace_bytes(int B, int T);
int be_potrf_batched(be_ctx* ctx, const double* A, int B, int T, double* L, int* info,
                     void* workspace, size_t workspace_bytes);

/* ---- a1 L1: posterior GPDTW1D.fit converges to for fixed kernel hyper-parameters -------
 * (models.py:185-220 at the natural-gradient fixed point).
 * Outputs: mu [B,T]; var_diag [B,T] = diag(cov); cov [B,T,T] (may be NULL);
 * scale_tri [B,T,T] = chol(cov), data.py:38-39 (may be NULL);
 * mvn_stats [B,4] = (|a|^2, a.b, |b|^2, sum log diag L) with a = L^-1 1, b = L^-1 mu --
 * everything LogLikelihoodWeight needs from a member (weights.py:97-100);
 * info_fit [B] (Cholesky of K + D + jitter I), info_dist [B] (Cholesky of cov). */
size_t be_gp_posterior_workspace_bytes(int B, int T, int R);
int be_gp_posterior(be_ctx* ctx, const double* X, const double* y_mean, const double* y_var,
                    const double* variance, const double* lengthscale, double jitter,
                    int B, int T, int R,
                    double* mu, double* var_diag, double* cov, double* scale_tri,
                    double* mvn_stats, int* info_fit, int* info_dist,
                    void* workspace, size_t workspace_bytes);

/* ---- a1 L2: the training loop GPDTW1D.fit actually runs, models.py:185-220 ----------------
 * Whitened VGP, Matern-3/2 on X, heteroskedastic Gaussian likelihood (a2, models.py:142-149;
 * its variational expectation enters through the closed-form natural-gradient target and the
 * analytic ELBO gradient).  n_iters times: NaturalGradient(gamma) step on (q_mu, q_sqrt)
 * (models.py:209), then -- if train_hypers -- one Adam(lr) step on the two softplus-constrained
 * kernel parameters (models.py:210; TF defaults beta1 .9, beta2 .999, eps 1e-7).  From GPflow's
 * initial state q_mu = 0, q_sqrt = I; variance / lengthscale [B] hold the initial values on entry
 * (GPflow: 1, 1) and the trained values on return.  Then predict_f(X, full_cov=True) (:217) and
 * cov += diag(y_var) (:220).  One iteration is captured in a CUDA graph and replayed.
 * Outputs as be_gp_posterior; cov [B,T,T] is required here. */
size_t be_vgp_fit_workspace_bytes(int B, int T, int R);
int be_vgp_fit(be_ctx* ctx, const double* X, const double* y_mean, const double* y_var, int B, int T, int R,
               int n_iters, double gamma, double lr, int train_hypers, double jitter,
               double* variance, double* lengthscale,
               double* mu, double* var_diag, double* cov, double* scale_tri, double* mvn_stats,
               int* info_fit, int* info_dist, void* workspace, size_t workspace_bytes);

/* be_gp_posterior_factored: the same fixed-theta posterior (a1 + a3) for callers that need only what
 * LogLikelihoodWeight (weights.py:97-100) and Barycentre (ensemble_scheme.py:63-65) consume -- mu [B,T],
 * var_diag [B,T] = diag(cov), mvn_stats [B,4] -- and not the dense cov / scale_tri.  The covariance
 * cov = E' - E M^-1 E (E = diag(y_var) + jitter I, E' = diag(y_var) + E, M = K + E) is kept in factored form:
 * cov^-1 = E'^-1 + G N^-1 G and det cov = det N det E' / det M with G = E E'^-1, N = K + diag(E D / E')
 * (Woodbury), so the statistics come from chol(M), its triangular inverse and chol(N): T^3 tensor flops per
 * problem instead of 4/3 T^3.  Agreement with be_gp_posterior: ~1e-12 relative (tests/test_gpu_parity.py).
 * info_dist reports chol(N).  N carries no jitter: where y_var is exactly 0 at every time step (a single
 * realisation) N = K can be singular although cov is not -- info_dist != 0 says so; use be_gp_posterior there. */
size_t be_gp_posterior_factored_workspace_bytes(int B, int T, int R);
int be_gp_posterior_factored(be_ctx* ctx, const double* X, const double* y_mean, const double* y_var,
                             const double* variance, const double* lengthscale, double jitter,
                             int B, int T, int R, double* mu, double* var_diag, double* mvn_stats,
                             int* info_fit, int* info_dist, void* workspace, size_t workspace_bytes);

/* ---- a3: Distribution(mu, cov, MultivariateNormalFullCovariance), data.py:38-39 ---------
 * from an arbitrary covariance: scale_tri, diag and the log-prob statistics. */
size_t be_mvn_from_cov_workspace_bytes(int B, int T);
int be_mvn_from_cov(be_ctx* ctx, const double* mu, const double* cov, int B, int T,
                    double* scale_tri, double* var_diag, double* mvn_stats, int* info,
                    void* workspace, size_t workspace_bytes);

/* ---- a4: LogLikelihoodWeight._compute, ensembles/weights.py:87-123 ----------------------
 * MVN branch (:97-100, quirk Q-LL): ll[c,m,r,i] = log N(obs[c,r,i] * 1_T | mu, Sigma).
 * obs [C,Ro,T]; mvn_stats [C*M,4].  lls_mean [C,M,T] = mean over r (:103-104);
 * lls_exp = exp(c * lls_mean) (:107, no max-subtraction, Q-EXP);
 * weights [C,M,T] = lls_exp / sum_m lls_exp (:122-123; 0/0 -> NaN kept).
 * lls_mean / lls_exp may be NULL. */
int be_loglik_weights_mvn(be_ctx* ctx, const double* mvn_stats, const double* obs,
                          int C, int M, int Ro, int T, double standardisation_constant,
                          double* weights, double* lls_exp, double* lls_mean);
/* per-realisation log-probs of the same branch, ll [C,M,Ro,T] (distribution.log_prob, :98-100) */
int be_mvn_constvec_logprob(be_ctx* ctx, const double* mvn_stats, const double* obs,
                            int C, int M, int Ro, int T, double* ll);
/* distrax MultivariateNormalTri.log_prob for N GENERAL vectors x [N,T] of one member (the NLL metric of
 * ensembles/utils.py:139 calls it with the held-out realisations): z = L^-1 (x - mu) by forward substitution
 * against the stored dense factor scale_tri [T,T] (data.py:38-39), sum_log_diag = mvn_stats[3]; ll [N]. */
int be_mvn_log_prob(be_ctx* ctx, const double* mu, const double* scale_tri, const double* x,
                    int T, int N, double sum_log_diag, double* ll);
/* dx.Normal branch (:95-96): elementwise log N(x | loc, scale) -- 2nd argument is a SCALE
 * (quirk Q-SCALE).  n elements. */
int be_normal_logprob(be_ctx* ctx, const double* loc, const double* scale, const double* x,
                      size_t n, double* ll);
/* Normal branch weights: loc/scale [C,M,N], obs [C,Ro,N] -> weights [C,M,N] */
int be_loglik_weights_normal(be_ctx* ctx, const double* loc, const double* scale,
                             const double* obs, int C, int M, int Ro, int N,
                             double standardisation_constant,
                             double* weights, double* lls_exp, double* lls_mean);
/* mean over the time axis skipping NaN (xarray .mean('time'), ensembles/utils.py:111),
 * broadcast back over time (utils.py:133): w [C,M,T] -> w_bar [C,M,T] */
int be_weights_time_mean(be_ctx* ctx, const double* weights, int C, int M, int T, double* w_bar);

/* member-sharded normalisation (weights.py:122-123 when the sum over models spans ranks):
 * weights [C,M_local,T] = lls_exp [C,M_local,T] / total [C,T] */
int be_weights_normalise(be_ctx* ctx, const double* lls_exp, const double* total, int C, int M, int T,
                         double* weights);

/* ---- a5/a6: Barycentre._compute + gaussian_barycentre ----------------------------------
 * ensembles/ensemble_scheme.py:54-72 and ensembles/wasserstein.py:61-100 (signed stop
 * rule, quirk Q-BARY).  means, variances, weights [C,M,N] -> mu, sigma [C,N]; iters [C,N]
 * (may be NULL) = iterations taken, > max_iters means "not converged" (the reference
 * warns, :94-97).  std = sqrt(variance) is taken inside (ensemble_scheme.py:65). */
int be_barycentre_1d(be_ctx* ctx, const double* means, const double* variances,
                     const double* weights, int C, int M, int N,
                     double tolerance, double init_var, int max_iters,
                     double* mu, double* sigma, int* iters);
/* member-sharded form (multi-GPU): partial sums over the local members
 *   partial [3,C,N] = (sum_m w, sum_m w mu, sum_m w sqrt(var)) for whatever weights are passed:
 * un-normalised w~ to obtain the normaliser (slot 0, all-reduced first, then be_weights_normalise),
 * then the normalised weights for slots 1-2; all-reduce over ranks (NCCL);
 * be_barycentre_1d_finish divides slots 1-2 by slot 0 and iterates. */
int be_barycentre_1d_partial(be_ctx* ctx, const double* means, const double* variances,
                             const double* lls_exp, int C, int M_local, int N, double* partial);
int be_barycentre_1d_finish(be_ctx* ctx, const double* partial, int C, int N,
                            double tolerance, double init_var, int max_iters,
                            double* mu, double* sigma, int* iters);

/* ---- a7: sqrtm, ensembles/wasserstein.py:10-13 --------------------------------------------
 * The reference computes U diag(sqrt(s)) V^H from jnp.linalg.svd; for SYMMETRIC POSITIVE DEFINITE
 * input -- all the path ever passes (covariances and S^1/2 Sigma S^1/2 products) -- that is the
 * principal square root, computed here by the scaled Denman-Beavers iteration on the FP64
 * tensor cores.  A [B,T,T] symmetric (both triangles read) -> sqrt_out [B,T,T];
 * inv_sqrt_out [B,T,T] (may be NULL) = A^-1/2.  Iterates until the relative Frobenius change of
 * the iterate is < tol (1e-10 recommended) or max_iters; *iters_host (HOST int, may be NULL)
 * receives the iteration count.  info [B] (device): k > 0 = an iterate was not positive definite
 * at leading minor k (input not SPD).  Synchronises the ctx stream once per iteration. */
size_t be_sqrtm_psd_workspace_bytes(int B, int T);
int be_sqrtm_psd(be_ctx* ctx, const double* A, int B, int T, double tol, int max_iters,
                 double* sqrt_out, double* inv_sqrt_out, int* iters_host, int* info,
                 void* workspace, size_t workspace_bytes);

/* ---- a8: gaussian_w2_distance_distrax, ensembles/wasserstein.py:21-47 ----------------------
 * w2[p] = |mu1 - mu2|_2 + tr(S1 + S2 - 2 sqrtm(sqrtm(S1) S2 sqrtm(S1)))   -- the location term
 * is NOT squared (quirk Q-W2, :40,45).  mu* [P,T], sigma* [P,T,T], w2 [P], info [P].
 * be_w2_distance_diag is the full_cov=False branch (:36-39): var* [P,T] are the variances the
 * reference puts on a diagonal, for which sqrtm is elementwise. */
size_t be_w2_distance_workspace_bytes(int P, int T);
int be_w2_distance(be_ctx* ctx, const double* mu1, const double* sigma1, const double* mu2,
                   const double* sigma2, int P, int T, double sqrtm_tol, int sqrtm_max_iters,
                   double* w2, int* info, void* workspace, size_t workspace_bytes);
int be_w2_distance_diag(be_ctx* ctx, const double* mu1, const double* var1, const double* mu2,
                        const double* var2, int P, int T, double* w2);

/* ---- a9: full-covariance Gaussian W2 barycentre (BASELINE config 5) -------------------------
 * The reference has no such code; this is the matrix generalisation of wasserstein.py:61-100
 * with wasserstein.py:10-13 as the square root, as DEFINED by oracle/reference_path.py
 * (fullcov_barycentre):  S0 = init_var I;  S <- sum_m w_m (S^1/2 Sigma_m S^1/2)^1/2;  signed stop
 * rule tr(S_new - S)/T < tolerance (reduces exactly to wasserstein.py:88 at T = 1);
 * mu = sum_m w_m mu_m (:98).  mus [C,M,T], sigmas [C,M,T,T], weights [C,M] -> mu [C,T],
 * S_out [C,T,T]; iters_host [C] (HOST, may be NULL) iterations taken, > max_iters = not
 * converged (the reference only warns); info [C*M] (device) as be_sqrtm_psd.
 * Synchronises the ctx stream inside. */
size_t be_barycentre_fullcov_workspace_bytes(int C, int M, int T);
int be_barycentre_fullcov(be_ctx* ctx, const double* mus, const double* sigmas, const double* weights,
                          int C, int M, int T, double tolerance, double init_var, int max_iters,
                          double sqrtm_tol, int sqrtm_max_iters,
                          double* mu, double* S_out, int* iters_host, int* info,
                          void* workspace, size_t workspace_bytes);

/* ---- SURVEY 8f "next" rows 2-3 (callers either side of the path) --------------------------
 * CRPSWeight._compute, ensembles/weights.py:444-515: crps_mean[c,m,n] = mean over obs realisations of
 * properscoring.crps_gaussian(obs, loc, scale) (:469-471); weights = (1/crps) normalised over
 * models (:507-511).  loc, scale [C,M,N] -- scale is the distribution's stddev(), which for the
 * dx.Normal(mean, variance) built at :497 is the member's VARIANCE (quirk Q-SCALE); obs [C,Ro,N];
 * weights [C,M,N]; crps_mean [C,M,N] may be NULL. */
int be_crps_weights(be_ctx* ctx, const double* loc, const double* scale, const double* obs,
                    int C, int M, int Ro, int N, double* weights, double* crps_mean);
/* KSDWeight._compute, ensembles/weights.py:336-441: ksd[c,m,n] = IMQ kernel Stein discrepancy (k_0_fun :360-375,
 * imq_KSD :380-394, c = 1, beta = -1/2) of the Ro observation samples at point n against
 * dx.Normal(loc, scale) (:417; scale = the member's VARIANCE, Q-SCALE), with grad log p = -(x - loc) / scale^2
 * (:419); weights = (1 / ksd) normalised over models (:434-438).  Layouts as be_crps_weights; ksd may be NULL. */
int be_ksd_weights(be_ctx* ctx, const double* loc, const double* scale, const double* obs,
                   int C, int M, int Ro, int N, double* weights, double* ksd);
/* ModelSimilarityWeight._compute, ensembles/weights.py:214-333.
 * be_w2_collapse: w2 [C,M,M,N] pairwise distances (from be_w2_distance / be_w2_distance_diag)
 *   -> nanmean over the second model (:259,296,321), normalised over models (:331): weights [C,M,N]
 *   (mode "single": N = 1; mode "spatial": N = lat * lon).
 * be_similarity_weights_pointwise: mode "temporal" (:302-325) in one pass -- per point the
 *   1-dimensional full_cov=False W2 of wasserstein.py:36-45 between every pair of members;
 *   mean, var [C,M,N] (var = the distributions' variance()); w2_out [C,M,M,N] may be NULL. */
int be_w2_collapse(be_ctx* ctx, const double* w2, int C, int M, int N, double* weights);
int be_similarity_weights_pointwise(be_ctx* ctx, const double* mean, const double* var,
                                    int C, int M, int N, double* weights, double* w2_out);

/* ---- SURVEY 8f "next" row 1: DTW barycentre averaging (the step that produces y_mean) ----------
 * be_dtw_barycenter_averaging_subgradient: tslearn 0.5.1.0 dtw_barycenter_averaging_subgradient as
 *   called at ensembles/models.py:176-178 (max_iter=50, tol=1e-3) and :251-253, batched over B
 *   independent (cell, member) problems.  X [B,R,T] realisations; init_barycenter [B,T] or NULL
 *   (= the mean over realisations, tslearn's _init_avg); barycenter [B,T] out; n_iter [B] (device,
 *   may be NULL) iterations run per problem; cost [B] (device, may be NULL) last cost evaluated.
 *   weights = None, barycenter_size = None, metric_params = None (the reference passes none).
 *   T <= 4096 (BE_ERR_UNSUPPORTED above).  Synchronises the ctx stream once per four iterations.
 * be_perform_dba: the reference's own NumPy DBA, ensembles/dtwa.py:6-20 (medoid initialisation
 *   :23-37, n_iterations of DBA_update :84-141), R <= 50 series of equal length; center [B,T] out,
 *   medoid [B] (device int, may be NULL).
 * be_dtw_squared: ensembles/dtwa.py:48-75 squared_DTW for P independent pairs A[p], X[p] ([P,T]). */
size_t be_dtw_dba_workspace_bytes(int B, int R, int T);
int be_dtw_barycenter_averaging_subgradient(be_ctx* ctx, const double* X, int B, int R, int T, int max_iter,
                                            double initial_step_size, double final_step_size, double tol,
                                            const double* init_barycenter, double* barycenter, int* n_iter,
                                            double* cost, void* workspace, size_t workspace_bytes);
int be_perform_dba(be_ctx* ctx, const double* X, int B, int R, int T, int n_iterations, double* center,
                   int* medoid, void* workspace, size_t workspace_bytes);
int be_dtw_squared(be_ctx* ctx, const double* A, const double* X, int P, int T, double* sqcost);

/* ---- f-4 (next row): the SVGP stage of GPDTW3D.fit, ensembles/models.py:357-411 --------------------------------
 * X [N,D] = (x, y, z, t_cont, R realisation columns), D = 4 + R (models.py:270-319); Y [N,2] = (DTW mean, variance).
 * Kernel = Matern32(active_dims=[3]) + Matern32([0,1]) + Matern32([2]) + Matern32([4..D)) (:358-364): variances /
 * lengthscales [4] in that order (in: initial values, GPflow's are 1; out: trained).  Z [M,D]: inducing inputs (in:
 * linspace(min X, max X, M), :370; out: trained -- they are trainable in GPflow).  SVGP(whiten=True, q_mu = 0,
 * q_sqrt = I, num_data=None: the ELBO is not rescaled) with the heteroskedastic Gaussian likelihood (:142-149).
 * Step s (n_steps = n_optim_nits * (N / minibatch_size), :393): NaturalGradient(gamma) on minibatch
 * batch_idx[2s] (:390), then, if train_hypers, Adam(lr) on the kernel parameters and Z on minibatch batch_idx[2s+1]
 * (:391; the loss closure draws a new batch per call).  batch_idx [2*n_steps, minibatch_size] int64 on the device: the
 * reference shuffles unseeded, so the caller supplies the order (oracle/svgp.py:batch_indices documents one).
 * Then predict_f(X, full_cov=False) in chunks of predict_chunk points (:408): mu [N], var [N] = fvar + Y[:,1] (:411).
 * q_mu [M], q_sqrt [M,M] (lower) are returned for inspection.  info [1]: LAPACK-style report of chol(Kuu). */
size_t be_svgp_fit_workspace_bytes(int N, int D, int M, int minibatch_size, int predict_chunk);
int be_svgp_fit(be_ctx* ctx, const double* X, const double* Y, int N, int D, int M, int minibatch_size,
                const long long* batch_idx, int n_steps, double gamma, double lr, int train_hypers, double jitter,
                int predict_chunk, double* Z, double* variances, double* lengthscales, double* q_mu, double* q_sqrt,
                double* mu, double* var, int* info, void* workspace, size_t workspace_bytes);

#ifdef __cplusplus
}
#endif
#endif /* BE_B200_H */
