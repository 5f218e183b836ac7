/*
 * ddw.h -- C ABI of libddw_b200.so: the B200 (sm_100a) replacement for the
 * third-party solver stack that deepdow's allocation layers call today.
 *
 * deepdow has no FFI of its own; what these entry points replace is the
 * cvxpylayers autograd Function invoked at
 *     deepdow/layers/allocate.py:273   NumericalMarkowitz   (CvxpyLayer built at :236-238)
 *     deepdow/layers/allocate.py:595   SparsemaxAllocator   (CvxpyLayer built at :560)
 *     deepdow/layers/allocate.py:513   SoftmaxAllocator, formulation="variational" (:475)
 *     deepdow/layers/allocate.py:691   NumericalRiskBudgeting (CvxpyLayer built at :669-671)
 * and the per-sample torch loop of
 *     deepdow/layers/misc.py:107-119   CovarianceMatrix.forward
 *                        :121-166      compute_covariance, :168-201 compute_sqrt.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to a contiguous row-major array that the caller
 *     owns (torch allocates them); nothing is allocated or freed in here, no global state
 *   - matrix / vector buffers of the Markowitz and covariance entry points must be 16-byte aligned
 *     (they are read and written with 128-bit accesses); torch allocations are
 *   - `dtype` selects the element type of all floating point buffers (DDW_F32 / DDW_F64)
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously
 *   - return value: 0 = enqueued, negative = argument / launch error (ddw_last_error())
 *   - per-problem solver status is written to `status` (int32, device)
 */
#ifndef DDW_H_
#define DDW_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDW_VERSION 102

enum ddw_dtype { DDW_F32 = 0, DDW_F64 = 1 };

/* shrinkage strategies of CovarianceMatrix (misc.py:58-68) */
enum ddw_shrinkage { DDW_SHRINK_NONE = 0, DDW_SHRINK_DIAGONAL = 1, DDW_SHRINK_IDENTITY = 2,
                     DDW_SHRINK_SCALED_IDENTITY = 3 };

/* per-problem status bits written by ddw_markowitz_fwd */
enum ddw_status {
    DDW_STATUS_OK = 0,
    DDW_STATUS_FALLBACK = 1,     /* primal-dual active set cycled; interior point fallback used  */
    DDW_STATUS_REGULARIZED = 2,  /* a non-positive pivot was lifted (Q singular on the free set) */
    DDW_STATUS_INEXACT = 4,      /* fallback did not reach an exact KKT point (result approximate) */
    DDW_STATUS_INFEASIBLE = 8    /* n_assets * max_weight < 1: the reference raises SolverError  */
};

/* error codes (return values) */
#define DDW_ERR_ARG (-1)
#define DDW_ERR_UNSUPPORTED (-2)
#define DDW_ERR_WORKSPACE (-3)
#define DDW_ERR_CUDA (-4)

int ddw_version(void);
/* message of the last error on the calling thread ("" if none) */
const char *ddw_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * NumericalMarkowitz -- replaces self.cvxpylayer(rets, gamma_sqrt_ * covmat_sqrt, alpha_abs)[0]
 * (allocate.py:240-273) including the gamma tiling (:268-270) and abs(alpha) (:271).
 *   minimise  w'Qw - rets'w,  Q = A'A + |alpha| I,  A[j,k] = gamma_sqrt[(b N^2 + j N + k) % B] * covmat_sqrt[b,j,k]
 *   s.t.      sum(w) = 1, 0 <= w <= max_weight
 * rets (B,N)  covmat_sqrt (B,N,N)  gamma_sqrt (B)  alpha (B)
 * w (B,N) out; state (B,N) int8 out: -1 at 0, +1 at max_weight, 0 free (saved for backward);
 * status (B) int32 out.  workspace: ddw_markowitz_workspace_bytes() bytes.  After the forward, the first
 * 16 bytes of the workspace hold int32 counters: [0] portfolios sent to the interior-point fallback,
 * [1] portfolios sent from the sparse-support kernel to the dense kernel, [2] portfolios left with
 * DDW_STATUS_INFEASIBLE or DDW_STATUS_INEXACT (the host layer raises SolverError when it is non-zero).
 * ------------------------------------------------------------------------------------------- */
size_t ddw_markowitz_workspace_bytes(int64_t B, int64_t N, int dtype);

int ddw_markowitz_fwd(const void *rets, const void *covmat_sqrt, const void *gamma_sqrt, const void *alpha,
                      double max_weight, int64_t B, int64_t N, int dtype,
                      void *w, int8_t *state, int32_t *status,
                      void *workspace, size_t workspace_bytes, void *stream);

/* backward of the above (replaces the diffcp adjoint inside cvxpylayers):
 * grad_w (B,N), w/state from the forward -> d_rets (B,N), d_covmat_sqrt (B,N,N), d_gamma_sqrt (B),
 * d_alpha (B).  d_gamma_sqrt may be NULL (skips the tiling scatter pass). */
int ddw_markowitz_bwd(const void *grad_w, const void *w, const int8_t *state,
                      const void *covmat_sqrt, const void *gamma_sqrt, const void *alpha,
                      double max_weight, int64_t B, int64_t N, int dtype,
                      void *d_rets, void *d_covmat_sqrt, void *d_gamma_sqrt, void *d_alpha,
                      void *workspace, size_t workspace_bytes, void *stream);

/* Forward / backward pair that keeps the factor of each converged reduced KKT system between the two calls (what autograd
 * does with saved tensors): the backward then runs substitutions only instead of forming Q_FF and factoring it again.
 * `saved` is an opaque DEVICE buffer of ddw_markowitz_saved_bytes() bytes, written by the forward and read by the backward
 * of the SAME batch; portfolios for which no factor was kept (small supports solved by the sparse-support kernel, interior
 * point fallbacks) take the regular backward path.  Results are those of ddw_markowitz_fwd / ddw_markowitz_bwd. */
size_t ddw_markowitz_saved_bytes(int64_t B, int64_t N, int dtype);
int ddw_markowitz_fwd_saved(const void *rets, const void *covmat_sqrt, const void *gamma_sqrt, const void *alpha,
                            double max_weight, int64_t B, int64_t N, int dtype,
                            void *w, int8_t *state, int32_t *status,
                            void *workspace, size_t workspace_bytes, void *saved, size_t saved_bytes, void *stream);
int ddw_markowitz_bwd_saved(const void *grad_w, const void *w, const int8_t *state,
                            const void *covmat_sqrt, const void *gamma_sqrt, const void *alpha,
                            double max_weight, int64_t B, int64_t N, int dtype,
                            void *d_rets, void *d_covmat_sqrt, void *d_gamma_sqrt, void *d_alpha,
                            void *workspace, size_t workspace_bytes, const void *saved, size_t saved_bytes, void *stream);

/* Fused CovarianceMatrix(sqrt=False) -> NumericalMarkowitz forward (BachelierNet's chain, deepdow/nn.py:170-171,189-191):
 * covmat_sqrt is never written to HBM -- each portfolio's (L, N) lookback tile x[b] is turned into its shrunk sample covariance
 * in shared memory (the code of ddw_covariance_fwd with do_sqrt = 0: same bits) and handed to the solver there, so the forward
 * reads L N + 2 N + 2 elements per portfolio instead of N^2 + 2 N + 2.  gamma_sqrt is applied with the reference's tiling as
 * in ddw_markowitz_fwd.  Returns DDW_ERR_UNSUPPORTED when the tile does not fit next to the solver (n_assets > 128 or a long
 * lookback): compose ddw_covariance_fwd and ddw_markowitz_fwd then.  covmat_scratch: (B, N, N) elements of UNINITIALISED
 * device memory: only a portfolio whose active set does not converge writes its covariance tile there, for the interior-point
 * fallback (which reads covmat_sqrt from HBM); with NULL such a portfolio is left DDW_STATUS_INEXACT instead.  `saved` (may be
 * NULL) as in ddw_markowitz_fwd_saved; the backward is ddw_covariance_fwd (recompute) + ddw_markowitz_bwd[_saved] +
 * ddw_covariance_bwd. */
int ddw_covariance_markowitz_fwd(const void *x, const void *coef, int strategy, int64_t L,
                                 const void *rets, const void *gamma_sqrt, const void *alpha,
                                 double max_weight, int64_t B, int64_t N, int dtype,
                                 void *w, int8_t *state, int32_t *status,
                                 void *workspace, size_t workspace_bytes, void *saved, size_t saved_bytes,
                                 void *covmat_scratch, void *stream);

/* Chunked launches of the two calls above: the same kernels on the `B` consecutive problems that start at
 * index `batch_offset` of a batch of `batch_total` problems.  Every pointer addresses the chunk's first
 * problem EXCEPT gamma_sqrt / d_gamma_sqrt, which address the whole batch's (batch_total) arrays: the
 * reference's repeat/view tiling (allocate.py:268-270) indexes gamma_sqrt by the flat element index of the
 * whole batch modulo batch_total.  With accumulate_gamma != 0 the chunk's contribution is added to
 * d_gamma_sqrt (the caller zeroes it once per batch) instead of overwriting it.
 * ddw_markowitz_fwd(B) == ddw_markowitz_fwd_chunk(B, batch_total = B, batch_offset = 0). */
int ddw_markowitz_fwd_chunk(const void *rets, const void *covmat_sqrt, const void *gamma_sqrt, const void *alpha,
                            double max_weight, int64_t B, int64_t N, int dtype, int64_t batch_total, int64_t batch_offset,
                            void *w, int8_t *state, int32_t *status,
                            void *workspace, size_t workspace_bytes, void *stream);
int ddw_markowitz_bwd_chunk(const void *grad_w, const void *w, const int8_t *state,
                            const void *covmat_sqrt, const void *gamma_sqrt, const void *alpha,
                            double max_weight, int64_t B, int64_t N, int dtype, int64_t batch_total, int64_t batch_offset,
                            void *d_rets, void *d_covmat_sqrt, void *d_gamma_sqrt, int accumulate_gamma, void *d_alpha,
                            void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * HOST-buffer entry point: what a CPU-resident caller of the reference layer binds.  The reference takes
 * CPU tensors, solves on host threads and returns CPU tensors (allocate.py:273); here the batch is streamed
 * through the GPU in chunks of `chunk` problems on three streams (host->device copies, solver kernels,
 * device->host copies), so the PCIe transfers in both directions overlap each other and the kernels.
 * All h_* pointers are HOST memory (page-locked for asynchronous copies; pageable memory works but
 * serialises); dev_scratch is DEVICE memory of ddw_markowitz_host_scratch_bytes() bytes owned by the caller.
 *   forward + backward : h_grad_w != NULL; writes h_w, h_status, h_d_rets, h_d_covmat_sqrt, h_d_alpha and
 *                        (if non-NULL) h_d_gamma_sqrt
 *   forward only       : h_grad_w == NULL; writes h_w, h_status (the h_d_* pointers are ignored)
 * Work is ordered after everything already enqueued on `stream` and `stream` is made to wait for the last
 * copy, so the call is asynchronous like the device entry points: synchronise `stream` (or record an event
 * on it) before reading the host outputs.  A pipeline handle owns three streams and their events, nothing else.
 * ------------------------------------------------------------------------------------------- */
typedef struct ddw_host_pipeline ddw_host_pipeline;
int ddw_host_pipeline_create(ddw_host_pipeline **out);
void ddw_host_pipeline_destroy(ddw_host_pipeline *pipe);
/* number of steps this handle served by replaying its cached CUDA graph: a step requested twice in a row with
 * identical arguments and page-locked host buffers is captured once and replayed with one launch afterwards */
int ddw_host_pipeline_graph_launches(const ddw_host_pipeline *pipe);
size_t ddw_markowitz_host_scratch_bytes(int64_t B, int64_t N, int dtype, int64_t chunk);
int ddw_markowitz_host_step(ddw_host_pipeline *pipe,
                            const void *h_rets, const void *h_covmat_sqrt, const void *h_gamma_sqrt,
                            const void *h_alpha, const void *h_grad_w,
                            double max_weight, int64_t B, int64_t N, int dtype, int64_t chunk,
                            void *h_w, int32_t *h_status,
                            void *h_d_rets, void *h_d_covmat_sqrt, void *h_d_gamma_sqrt, void *h_d_alpha,
                            void *dev_scratch, size_t dev_scratch_bytes, void *stream);

/* The same step with the gradient of covmat_sqrt returned in FACTORED form: d_covmat_sqrt is rank 2 per portfolio,
 *     d_covmat_sqrt[b, j, k] = -2 gamma_[b, j, k] * (p[b, j] w[b, k] + q[b, j] d_rets[b, k]),   p = A d, q = A w,
 * so h_factors (B, 2, N) = (p, q) carries it in 2 N numbers instead of N^2 (1.6 KB instead of 40 KB at n_assets = 100): the
 * device->host half of the transfer shrinks from 167 MB to 6.6 MB at B = 4096 and the caller forms the outer products where
 * (and if) it needs them.  gamma_ is the reference's tiling (allocate.py:268-270).  Everything else as above. */
int ddw_markowitz_host_step_factored(ddw_host_pipeline *pipe,
                                     const void *h_rets, const void *h_covmat_sqrt, const void *h_gamma_sqrt,
                                     const void *h_alpha, const void *h_grad_w,
                                     double max_weight, int64_t B, int64_t N, int dtype, int64_t chunk,
                                     void *h_w, int32_t *h_status,
                                     void *h_d_rets, void *h_factors, void *h_d_gamma_sqrt, void *h_d_alpha,
                                     void *dev_scratch, size_t dev_scratch_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * NumericalRiskBudgeting -- replaces self.cvxpylayer(covmat_sqrt, b)[0] (deepdow/layers/allocate.py:673-691; the
 * program is stated at :651-671):
 *   minimise  1/2 ||covmat_sqrt w||^2 - b' log(w)   s.t.  sum(w) = 1, 0 <= w <= max_weight
 * covmat_sqrt (B,N,N)  b (B,N) risk budgets (nonnegative; zeros are floored at the smallest normal number)
 * w (B,N) out; state (B,N) int8 out: +1 at max_weight, 0 free (saved for backward); status (B) int32 out: the
 * ddw_status bits in the low 16 bits, the number of Newton / active-set iterations used above them.
 * After the forward the first int32 of the workspace holds the number of portfolios left DDW_STATUS_INFEASIBLE or
 * DDW_STATUS_INEXACT.  The backward is the implicit differentiation of the KKT system on the free set:
 * grad_w (B,N) -> d_covmat_sqrt (B,N,N), d_b (B,N).
 * ------------------------------------------------------------------------------------------- */
size_t ddw_risk_budgeting_workspace_bytes(int64_t B, int64_t N, int dtype);
int ddw_risk_budgeting_fwd(const void *covmat_sqrt, const void *b, double max_weight, int64_t B, int64_t N, int dtype,
                           void *w, int8_t *state, int32_t *status, void *workspace, size_t workspace_bytes, void *stream);
int ddw_risk_budgeting_bwd(const void *grad_w, const void *w, const int8_t *state, const void *covmat_sqrt, const void *b,
                           double max_weight, int64_t B, int64_t N, int dtype, void *d_covmat_sqrt, void *d_b,
                           void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * SparsemaxAllocator -- replaces self.layer(x / temperature[:, None])[0] (allocate.py:586-595):
 * Euclidean projection of x/temperature onto {sum = 1, 0 <= w <= max_weight}.
 * x (B,N); temperature (B) or NULL (then temperature_scalar is used for every row); w (B,N) out.
 * ------------------------------------------------------------------------------------------- */
int ddw_capped_simplex_fwd(const void *x, const void *temperature, double temperature_scalar,
                           double max_weight, int64_t B, int64_t N, int dtype, void *w, void *stream);
/* d_x (B,N) out; d_temperature (B) out or NULL */
int ddw_capped_simplex_bwd(const void *grad_w, const void *w, const void *x, const void *temperature,
                           double temperature_scalar, double max_weight, int64_t B, int64_t N, int dtype,
                           void *d_x, void *d_temperature, void *stream);

/* ---------------------------------------------------------------------------------------------
 * SoftmaxAllocator(formulation="variational") -- replaces self.layer(inp)[0] (allocate.py:508-514):
 * argmin -x'w - sum entr(w)  s.t. sum = 1, w <= max_weight, i.e. w_i = min(max_weight, exp(x_i/t - nu)).
 * ------------------------------------------------------------------------------------------- */
int ddw_capped_softmax_fwd(const void *x, const void *temperature, double temperature_scalar,
                           double max_weight, int64_t B, int64_t N, int dtype, void *w, void *stream);
int ddw_capped_softmax_bwd(const void *grad_w, const void *w, const void *x, const void *temperature,
                           double temperature_scalar, double max_weight, int64_t B, int64_t N, int dtype,
                           void *d_x, void *d_temperature, void *stream);

/* ---------------------------------------------------------------------------------------------
 * MaximumReturn benchmark -- replaces self.optlayer(rets_estimate)[0] (deepdow/benchmarks.py:114-125, :160):
 * argmax x'w  s.t. sum = 1, 0 <= w <= max_weight.  The linear program's vertex solution: floor(1/max_weight) largest
 * entries at the cap, the next one takes the remainder; ties go to the smaller index.  Inference only (benchmarks are
 * not trainable, benchmarks.py:13-15).  DDW_ERR_ARG when N * max_weight < 1.
 * ------------------------------------------------------------------------------------------- */
int ddw_capped_argmax_fwd(const void *x, double max_weight, int64_t B, int64_t N, int dtype, void *w, void *stream);

/* ---------------------------------------------------------------------------------------------
 * CovarianceMatrix -- replaces the python loop of misc.py:107-119.
 * x (B,L,N) observations; coef (B) shrinkage coefficient per sample (ignored for DDW_SHRINK_NONE);
 * out (B,N,N): shrunk sample covariance, or its symmetric PSD square root when do_sqrt != 0.
 * When do_sqrt != 0, eigvec (B,N,N) and eigval (B,N) receive the eigen-decomposition of the shrunk
 * covariance (saved for backward; eigval holds sqrt|lambda| with sub-threshold values zeroed,
 * misc.py:185-199).  They may be NULL when no backward is needed; then workspace supplies scratch.
 * L == 1 is rejected by the host layer with ZeroDivisionError as in misc.py:143.
 * ------------------------------------------------------------------------------------------- */
size_t ddw_covariance_workspace_bytes(int64_t B, int64_t L, int64_t N, int dtype, int do_sqrt);

int ddw_covariance_fwd(const void *x, const void *coef, int strategy, int do_sqrt,
                       int64_t B, int64_t L, int64_t N, int dtype,
                       void *out, void *eigvec, void *eigval, int32_t *sweeps,
                       void *workspace, size_t workspace_bytes, void *stream);

/* grad_out (B,N,N) -> d_x (B,L,N), d_coef (B) (may be NULL). */
int ddw_covariance_bwd(const void *grad_out, const void *x, const void *coef, int strategy, int do_sqrt,
                       const void *eigvec, const void *eigval,
                       int64_t B, int64_t L, int64_t N, int dtype,
                       void *d_x, void *d_coef,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * CovarianceMatrix.compute_sqrt (misc.py:168-201) on given symmetric matrices m (B,N,N):
 * out = v sqrt|lambda| v' with the reference's rank threshold; eigvec/eigval as in ddw_covariance_fwd
 * (NULL when no backward is needed).  Workspace: ddw_covariance_workspace_bytes(B, 2, N, dtype, 1).
 * ------------------------------------------------------------------------------------------- */
int ddw_psd_sqrt_fwd(const void *m, int64_t B, int64_t N, int dtype, void *out, void *eigvec, void *eigval,
                     void *workspace, size_t workspace_bytes, void *stream);
int ddw_psd_sqrt_bwd(const void *grad_out, const void *eigvec, const void *eigval, int64_t B, int64_t N, int dtype,
                     void *d_m, void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DDW_H_ */
