/*
 * bbgpu.h -- C-ABI of libbbgpu.so: the B200 (sm_100a) implementation of bayesbridge's
 * CG-accelerated coefficient update and Polya-Gamma draw.
 *
 * Every entry point replaces one seam of the reference (OHDSI/bayes-bridge v0.2.6, paths
 * relative to the reference root); the cited lines are what a maintainer would re-bind:
 *
 *   seam 1  design matrix      bayesbridge/design_matrix/sparse_matrix.py:21-49   (construction)
 *                              bayesbridge/design_matrix/sparse_matrix.py:68-101  (dot)
 *                              bayesbridge/design_matrix/sparse_matrix.py:103-129 (Tdot)
 *                              bayesbridge/design_matrix/sparse_matrix.py:164-177 (fisher diag)
 *                              bayesbridge/design_matrix/dense_matrix.py:9-58     (dense twin)
 *   seam 2  CG sampler         bayesbridge/reg_coef_sampler/cg_sampler.py:20-94
 *                              (+ scipy.sparse.linalg.cg, the third-party loop it calls)
 *   seam 3  random variates    bayesbridge/random/random.py:37-41 -> polya_gamma.pyx:40-74,
 *                              tilted_stable.pyx:65-135
 *
 * Conventions: plain pointers and sizes only; all host buffers are caller-owned, C-contiguous,
 * fp64 / int32, and are only touched during the call (every call returns after its stream has
 * been synchronised).  Device memory is owned by the handles.  Every function returns 0 on
 * success, non-zero on failure; bb_last_error() then holds the message.  A handle is not
 * thread-safe.  One process drives one GPU; multi-GPU = one process per GPU + bb_comm_*.
 */
#ifndef BBGPU_H
#define BBGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bb_ctx bb_ctx;   /* device + stream + (optional) communicator            */
typedef struct bb_mat bb_mat;   /* one row shard of the design matrix + its workspaces  */

enum { BB_OK = 0, BB_ERR_CUDA = 1, BB_ERR_ARG = 2, BB_ERR_NCCL = 3, BB_ERR_STATE = 4 };

/* noise source of the CG right-hand side (cg_sampler.py:61-67) */
enum { BB_NOISE_INJECT = 0,   /* eps1[n_local], eps2[P] supplied by the caller (parity mode)   */
       BB_NOISE_PHILOX = 1 }; /* generated on device, Philox4x32-10 keyed by (seed, offset, global row) */

/* spmv kernel variants (bb_set_option "spmv_stage", read when a matrix is uploaded):
 * gather vector staged in shared memory (slab formats) or read through L2 (single slab) */
enum { BB_SPMV_STAGE_NONE = 0, BB_SPMV_STAGE_BOTH = 1, BB_SPMV_STAGE_DOT = 2, BB_SPMV_STAGE_TDOT = 3 };

/* ---- library ---------------------------------------------------------------------------- */
const char* bb_last_error(void);
int  bb_version(void);
int  bb_device_count(int* count);

/* ---- context ---------------------------------------------------------------------------- */
int  bb_init(int device, bb_ctx** out);
int  bb_destroy(bb_ctx* ctx);
int  bb_set_option(bb_ctx* ctx, const char* name, int64_t value);
int  bb_get_option(bb_ctx* ctx, const char* name, int64_t* value);
/* kernel launches issued by this library since bb_init / the last reset */
int  bb_get_launch_count(bb_ctx* ctx, int64_t* launches);
int  bb_reset_launch_count(bb_ctx* ctx);
/* device milliseconds (CUDA events on the library stream) spent inside the entry points below */
int  bb_get_device_ms(bb_ctx* ctx, double* ms);
int  bb_reset_device_ms(bb_ctx* ctx);
int  bb_sync(bb_ctx* ctx);

/* ---- communicator: row-sharding, one process per GPU (new) ----------- */
int  bb_comm_unique_id(const char* nccl_lib_path, char* id_out_128);
int  bb_comm_init(bb_ctx* ctx, const char* nccl_lib_path, int nranks, int rank, const char* id_128);
/* communicator WITHOUT NCCL: every exchange goes through the peer-memory all-reduce below (vectors longer than its
 * capacity are reduced in chunks).  NCCL refuses two ranks on one device; this form does not, which is how the
 * row-sharded path is tested on a single GPU (tests/test_gpu_multi.py).  Call bb_comm_p2p_export/attach next. */
int  bb_comm_init_local(bb_ctx* ctx, int nranks, int rank);
int  bb_comm_allreduce_host(bb_ctx* ctx, double* buf, int64_t count);  /* in-place sum, host buffer */
/* one-shot all-reduce over NVLink peer memory (replaces ncclAllReduce for vectors of <= capacity doubles):
 * every rank exports its exchange buffer (64-byte CUDA IPC handle), the handles are gathered by the caller
 * (any transport) and attached; option "allreduce_p2p" = 0 switches back to NCCL. */
int  bb_comm_p2p_export(bb_ctx* ctx, int64_t capacity, char* handle_out_64);
int  bb_comm_p2p_attach(bb_ctx* ctx, const char* handles_nranks_x_64);
int  bb_comm_p2p_status(bb_ctx* ctx, int* ready, int* error);

/* ---- design matrix (seam 1) ------------------------------------------------------------- */
/* CSR shard [n_local x p]; data==NULL => pattern-only (all stored values are 1.0).
 * column_offset==NULL => not centred.  row_offset/n_global place the shard in the global matrix. */
int  bb_csr_upload(bb_ctx* ctx, int64_t n_local, int64_t p, int64_t nnz,
                   const int32_t* indptr, const int32_t* indices, const double* data,
                   const double* column_offset, int add_intercept,
                   int64_t row_offset, int64_t n_global, bb_mat** out);
/* dense shard, row-major [n_local x p] WITHOUT intercept column / centring (applied implicitly) */
int  bb_dense_upload(bb_ctx* ctx, int64_t n_local, int64_t p, const double* X,
                     const double* column_offset, int add_intercept,
                     int64_t row_offset, int64_t n_global, bb_mat** out);
int  bb_mat_free(bb_mat* mat);
int  bb_mat_info(bb_mat* mat, int64_t* n_local, int64_t* P, int64_t* nnz, int* is_sparse, int* is_binary);
/* bit-exactness hooks: what the device holds, copied back */
int  bb_mat_export_csr(bb_mat* mat, int32_t* indptr, int32_t* indices, double* data);
int  bb_mat_export_csc(bb_mat* mat, int32_t* indptr, int32_t* indices, double* data);

/* y[n_local] = X v  (with intercept + centring algebra);  v has P = p + add_intercept entries */
int  bb_dot(bb_mat* mat, const double* v, double* out);
/* t[P] = X' w, summed over all shards when a communicator is attached */
int  bb_tdot(bb_mat* mat, const double* w, double* out);
/* diag(X' diag(weight) X)[P], summed over shards */
int  bb_fisher_diag(bb_mat* mat, const double* weight, double* out);
/* X' diag(weight) X, P x P row-major (symmetric), summed over shards: compute_fisher_info(weight, diag_only=False),
 * design_matrix/dense_matrix.py:54-58, sparse_matrix.py:131-162.  fp64 tensor-core (mma.sync m8n8k4) tile kernel;
 * a sparse design is densified on the device first.  device_ms (nullable): milliseconds of the X'WX kernel alone. */
int  bb_fisher_full(bb_mat* mat, const double* weight, double* out, double* device_ms);
/* The direct ("cholesky") Gaussian draw of reg_coef_sampler/direct_gaussian_sampler.py:4-44 (generate_gaussian_with_weight):
 * coef ~ N(Sigma z, Sigma), Sigma^-1 = X' diag(omega) X + diag(prior_prec_sqrt)^2, Jacobi-scaled, upper Cholesky factor,
 * with the standard normal vector `gaussian_vec[P]` supplied by the caller (the reference draws it from numpy's global
 * stream).  omega: host pointer or NULL for the resident precisions.  stats (nullable): [ms X'WX, ms factorisation]. */
int  bb_cholesky_sample(bb_mat* mat, const double* omega, const double* prior_prec_sqrt, const double* z,
                        const double* gaussian_vec, double* coef_out, double* stats);

/* Column sums and sums of squares over the LOCAL rows, from the resident CSC image: the moments the constructor needs for
 * remove_intercept_indicator and for the centring offsets (design_matrix/abstract_matrix.py:93-107, sparse_matrix.py:38-45)
 * without two host passes over the matrix. */
int  bb_column_moments(bb_mat* mat, double* sum_out, double* sumsq_out);
/* (re)sets the centring offsets (column means) of a resident design; NULL = not centred */
int  bb_set_column_offset(bb_mat* mat, const double* offset);

/* ---- resident observation-side vectors (avoid n-length PCIe traffic per Gibbs iteration) - */
/* logit: n_trial, n_success;  linear: n_trial==NULL, n_success = y */
int  bb_set_outcome(bb_mat* mat, const double* n_trial, const double* n_success);
int  bb_set_obs_prec(bb_mat* mat, const double* omega);          /* H2D, n_local            */
int  bb_set_obs_prec_scalar(bb_mat* mat, double omega);          /* linear model: omega*1_n */
int  bb_get_obs_prec(bb_mat* mat, double* omega_out);            /* D2H                     */
int  bb_get_linear_predictor(bb_mat* mat, double* eta_out);      /* D2H of the last X beta  */

/* ---- CG sampler (seam 2) ---------------------------------------------------------------- */
/* Draws beta ~ N(Phi^-1 z, Phi^-1), Phi = X' Omega X + diag(prior_prec_sqrt)^2, by running
 * scipy.sparse.linalg.cg's recurrences on the system pre-scaled by precond_scale.
 *   omega           NULL => use the resident obs_prec
 *   z               NULL => z = X'(omega .* y_gaussian) from the resident outcome
 *                           (logit: X' (n_success - n_trial/2); linear: omega * X' y)
 *   x0              initial guess for beta (un-scaled), P
 *   precond_scale   s in cg_sampler.py:123-138, P
 *   eps1, eps2      BB_NOISE_INJECT only
 * out: coef[P]; n_iter = completed CG iterations; info = 0 converged / maxiter otherwise.
 * stats (may be NULL): [0]=||b||, [1]=final ||r||, [2]=device ms of the call */
int  bb_cg_sample(bb_mat* mat, const double* omega, const double* prior_prec_sqrt,
                  const double* z, const double* x0, const double* precond_scale,
                  double atol, int maxiter, int noise_mode,
                  const double* eps1, const double* eps2, uint64_t seed, uint64_t offset,
                  double* coef_out, int* n_iter, int* info, double* stats);

/* ---- random variates (seam 3) ----------------------------------------------------------- */
/* omega_i ~ PG(shape_i, tilt_i); index_offset = global index of element 0 (sharding-invariant streams) */
int  bb_pg_sample(bb_ctx* ctx, int64_t n, const int32_t* shape, const double* tilt,
                  uint64_t seed, uint64_t offset, int64_t index_offset, double* out);
/* fused Gibbs step: eta = X coef; omega ~ PG(n_trial, eta) kept resident; loglik = sum(n_success*eta
 * - n_trial*log(1+e^eta)) over all shards.  omega_out may be NULL; coef may be NULL = the coefficients of the last
 * CG draw, still resident on the device. */
int  bb_pg_from_coef(bb_mat* mat, const double* coef, uint64_t seed, uint64_t offset,
                     double* omega_out, double* loglik);
/* linear model: sum of squared residuals ||y - X coef||^2 over all shards */
int  bb_linear_rss(bb_mat* mat, const double* coef, double* rss);
/* exponentially tilted stable draws (tilted_stable.pyx:65-135); char_exp scalar */
int  bb_tilted_stable_sample(bb_ctx* ctx, int64_t n, double char_exp, const double* tilt,
                             uint64_t seed, uint64_t offset, int64_t index_offset, double* out);
/* standard normals from the same Philox streams the CG sampler uses (stream 0: eps1, 1: eps2) */
int  bb_philox_normal(bb_ctx* ctx, int64_t n, int stream, uint64_t seed, uint64_t offset,
                      int64_t index_offset, double* out);

/* ---- device-resident P-side Gibbs state (reference: reg_coef_sampler.py:60-103,
 * reg_coef_posterior_summarizer.py:12-124, bayesbridge.py:458-478) -------------------------------------------
 * The local scales and the running summaries of the prior-scaled coefficients live on the device; the vectors the
 * CG sampler needs are formed there, so that per Gibbs iteration only coef (out) and a few scalars cross PCIe. */
int  bb_state_init(bb_mat* mat, int n_unshrunk, const double* prior_sd_unshrunk, double slab_size);
int  bb_state_set(bb_mat* mat, const double* lscale, const double* mean, const double* square, int64_t n_averaged);
int  bb_state_get(bb_mat* mat, double* lscale, double* mean, double* square, int64_t* n_averaged);
/* beta | omega, tau, lambda with device noise; sums_out = {sum|beta_shrunk|^bridge_exp, #nonzero shrunk,
 * sum (beta/slab)^2, sum (beta_unshrunk/prior_sd)^2}; the summaries are updated with the new draw */
int  bb_cg_sample_resident(bb_mat* mat, const double* omega, double gscale, double bridge_exp,
                           double atol, int maxiter, uint64_t seed, uint64_t offset,
                           double* coef_out, int* n_iter, int* info, double* sums_out);
/* lambda | tau, beta from the coefficients of the last bb_cg_sample_resident; counts_out = {#tilt<=0, #zeros, #inf};
 * lscale_out may be NULL */
int  bb_local_scale_resident(bb_mat* mat, double gscale, double char_exp, uint64_t seed, uint64_t offset,
                             int* counts_out, double* lscale_out);

/* Log-likelihood and gradient with the outcome resident (bb_set_outcome): only P-length vectors cross PCIe.
 * logit : model/logistic_model.py:49-55   ll = sum n_success eta - n_trial log(1+e^eta), grad = X'(n_success - n_trial sigmoid(eta))
 * linear: model/linear_model.py:13-24     ll = -obs_prec/2 ||y - X coef||^2 (the n/2 log obs_prec term is added by the caller),
 *                                         grad = obs_prec X'(y - X coef)
 * Used by the L-BFGS mode search that initialises a chain (reg_coef_sampler.py:281-327).  Sums run over all shards. */
int  bb_loglik_and_gradient(bb_mat* mat, const double* coef, double obs_prec, int loglik_only, double* loglik, double* grad);

/* ---- batched multi-chain form (BASELINE config 5; SURVEY section 8b "batched form with leading chain dimension") -------
 * C <= 16 independent chains on ONE dense design: the CG iterations of all chains run in lock-step, so that the two
 * products of an operator application become X V (n x C) and X'(Omega o U) (p x C) on the fp64 tensor cores and X is read
 * once for all chains.  Host arrays are chain-major: [C][n] for observation-side, [C][P] for coefficient-side vectors.
 * Every chain keeps the semantics of bb_cg_sample (reg_coef_sampler/cg_sampler.py:20-94 + scipy's cg): own omega, prior,
 * right-hand-side noise (injected, or Philox keyed by (seeds[c], offsets[c], global row)), own stopping decision. */
int  bb_batch_init(bb_mat* mat, int n_chains);
int  bb_batch_free(bb_mat* mat);
int  bb_batch_set_obs_prec(bb_mat* mat, const double* omega);            /* [C][n] */
int  bb_batch_get_obs_prec(bb_mat* mat, double* omega_out);              /* [C][n] */
int  bb_dot_batched(bb_mat* mat, const double* v, double* out);          /* out[c] = X v[c]   ([C][P] -> [C][n]) */
int  bb_tdot_batched(bb_mat* mat, const double* w, double* out);         /* out[c] = X' w[c]  ([C][n] -> [C][P]) */
int  bb_cg_sample_batched(bb_mat* mat, const double* omega, const double* prior_prec_sqrt, const double* z,
                          const double* x0, const double* precond_scale, double atol, int maxiter, int noise_mode,
                          const double* eps1, const double* eps2, const uint64_t* seeds, const uint64_t* offsets,
                          double* coef_out, int* n_iter, int* info);
/* omega_c | beta_c for every chain (bayesbridge.py:397-410, polya_gamma.pyx:40-216) + the logistic log-likelihoods;
 * coef: [C][P] or NULL for the coefficients of the last batched draw; the precisions stay resident for the next draw */
int  bb_pg_from_coef_batched(bb_mat* mat, const double* coef, const uint64_t* seeds, const uint64_t* offsets, double* loglik);

/* The mode search that initialises a chain (reg_coef_sampler/reg_coef_sampler.py:281-358: scipy L-BFGS-B without bounds,
 * maxcor 200, gtol 1e-6/sqrt(P), maxiter 250) run entirely on the device: minimises
 *   F(theta) = -loglik(scale . theta) + 1/2 sum prior_prec theta^2,  theta = coef / scale,
 * by L-BFGS (two-loop recursion in one kernel per iteration, strong-Wolfe line search), same stopping rules.
 * status: 0 gradient tolerance met, 1 relative decrease <= ftol, 2 maxiter reached, 3 line search failed. */
int  bb_mode_search(bb_mat* mat, const double* coef0, const double* scale, const double* prior_prec, double obs_prec,
                    int maxiter, double gtol, double ftol, int maxcor, double* coef_out,
                    int* n_iter, int* n_eval, int* status);

/* ---- timing ----------------------------------------------------------------------------- */
/* runs `reps` launches of one kernel class on resident data and returns mean device ms:
 * what = "dot" | "tdot" | "op" (one application of X' Omega X v, two products) | "spmv_dot" | "spmv_tdot" (the sparse
 * kernel alone) | "fused_op" (dense: the one-pass operator) | "exchange" (the all-reduce of a (p+1)-vector) */
int  bb_time_kernel(bb_mat* mat, const char* what, int reps, int flush_l2, double* ms_out);
/* profiling aid: per-CTA time line (ns of %globaltimer) of ONE launch of the sliced SpMV kernel, taken with a debug
 * instantiation of the kernel.  which = 0: the dot format, 1: the Tdot format.  out[cta * 40 + k]: k = 0 kernel entry,
 * 1 dependency wait passed, 2 first gather window staged, 3 CTA end, 4 number of sections, 8..39 end of each warp's strip.
 * capacity = words available in `out` (>= 40 * number of SMs).  (introspection, SURVEY section 8b "bb_get_timers") */
int  bb_spmv_timeline(bb_mat* mat, int which, int flush_l2, uint64_t* out, int64_t capacity, int* ncta_out);
/* measured fp64 tensor-core throughput (mma.sync m8n8k4 f64 issued back to back from registers), TFLOP/s: the roofline
 * denominator of the X'WX kernel (MEASURED_PEAKS.json carries no fp64 figure) */
int  bb_measure_fp64_mma(bb_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* BBGPU_H */
