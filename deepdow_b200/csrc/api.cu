// api.cu -- extern "C" surface of libddw_b200.so (declared in include/ddw.h)
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace ddw {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return DDW_ERR_CUDA;
    }
    return 0;
}

// implemented in markowitz_impl.cuh (instantiated in markowitz_{fwd,bwd}_{f32,f64}.cu) / rowproj.cu / covariance.cu
size_t markowitz_workspace_bytes(long long B, int N, int tsize);
size_t markowitz_saved_bytes(long long B, int N, int tsize);
template <typename T>
int markowitz_fwd_t(const void *, const void *, const void *, const void *, double, long long, int, void *, int8_t *,
                    int32_t *, void *, size_t, cudaStream_t, long long, long long, void *);
template <typename T>
int markowitz_bwd_t(const void *, const void *, const int8_t *, const void *, const void *, const void *, long long,
                    int, void *, void *, void *, void *, void *, size_t, cudaStream_t, long long, long long, int, const void *,
                    void *);
template <typename T>
int row_op_t(int, const void *, const void *, double, double, long long, int, const void *, const void *, void *,
             void *, void *, cudaStream_t);
template <typename T>
int markowitz_fused_fwd_t(const void *, const void *, int, int, const void *, const void *, const void *, double, long long, int,
                          void *, int8_t *, int32_t *, void *, size_t, cudaStream_t, void *, void *);
size_t risk_budgeting_workspace_bytes(long long B, int N, int tsize);      // risk_budgeting.cu
template <typename T>
int risk_budgeting_t(bool, const void *, const void *, const void *, const void *, const int8_t *, double, long long, int, void *,
                     int8_t *, int32_t *, void *, void *, void *, size_t, cudaStream_t);
size_t covariance_workspace_bytes(long long B, int L, int N, int tsize, int do_sqrt);
template <typename T>
int covariance_fwd_t(const void *, const void *, int, int, long long, int, int, void *, void *, void *, int32_t *,
                     void *, size_t, cudaStream_t);
template <typename T>
int covariance_bwd_t(const void *, const void *, const void *, int, int, const void *, const void *, long long, int,
                     int, void *, void *, void *, size_t, cudaStream_t);

static bool misaligned(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }
#define DDW_CHECK_ALIGNED(fn, ...)                                                              \
    do {                                                                                        \
        const void *ptrs_[] = {__VA_ARGS__};                                                    \
        for (const void *q_ : ptrs_)                                                            \
            if (misaligned(q_)) { set_error("%s: every buffer must be 16-byte aligned", fn); return DDW_ERR_ARG; } \
    } while (0)

static int check_dims(const char *fn, int64_t B, int64_t N, int64_t nmax, int dtype) {
    if (dtype != DDW_F32 && dtype != DDW_F64) { set_error("%s: dtype must be DDW_F32 or DDW_F64", fn); return DDW_ERR_ARG; }
    if (B < 0 || B >= (1LL << 31)) { set_error("%s: batch size %lld out of range", fn, (long long)B); return DDW_ERR_ARG; }
    if (N < 1 || N > nmax) { set_error("%s: n_assets=%lld outside [1, %lld]", fn, (long long)N, (long long)nmax); return DDW_ERR_UNSUPPORTED; }
    return 0;
}

}  // namespace ddw

using namespace ddw;

extern "C" {

int ddw_version(void) { return DDW_VERSION; }

const char *ddw_last_error(void) { return g_err; }

size_t ddw_markowitz_workspace_bytes(int64_t B, int64_t N, int dtype) {
    if (B <= 0 || N <= 0 || N > 1024) return 0;
    return markowitz_workspace_bytes(B, (int)N, dtype == DDW_F64 ? 8 : 4);
}

size_t ddw_markowitz_saved_bytes(int64_t B, int64_t N, int dtype) {
    if (B <= 0 || N <= 0 || N > 1024) return 0;
    return markowitz_saved_bytes(B, (int)N, dtype == DDW_F64 ? 8 : 4);
}

static int markowitz_fwd_impl(const void *rets, const void *covmat_sqrt, const void *gamma_sqrt, const void *alpha,
                              double max_weight, int64_t B, int64_t N, int dtype, int64_t batch_total, int64_t batch_offset,
                              void *w, int8_t *state, int32_t *status, void *workspace, size_t workspace_bytes,
                              void *saved, size_t saved_bytes, void *stream) {
    g_err[0] = 0;
    int rc = check_dims("ddw_markowitz_fwd", B, N, 1024, dtype);
    if (rc) return rc;
    if (B == 0) return 0;
    if (batch_offset < 0 || batch_offset + B > batch_total || batch_total >= (1LL << 31)) { set_error("ddw_markowitz_fwd: chunk [%lld, %lld) outside the batch of %lld", (long long)batch_offset, (long long)(batch_offset + B), (long long)batch_total); return DDW_ERR_ARG; }
    if (!rets || !covmat_sqrt || !gamma_sqrt || !alpha || !w || !state || !status) { set_error("ddw_markowitz_fwd: null pointer"); return DDW_ERR_ARG; }
    DDW_CHECK_ALIGNED("ddw_markowitz_fwd", rets, covmat_sqrt, gamma_sqrt, alpha, w, workspace, saved);
    if (saved && saved_bytes < ddw_markowitz_saved_bytes(B, N, dtype)) { set_error("ddw_markowitz_fwd_saved: saved buffer too small (%zu < %zu)", saved_bytes, ddw_markowitz_saved_bytes(B, N, dtype)); return DDW_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == DDW_F64
               ? markowitz_fwd_t<double>(rets, covmat_sqrt, gamma_sqrt, alpha, max_weight, B, (int)N, w, state, status, workspace, workspace_bytes, st, batch_total, batch_offset, saved)
               : markowitz_fwd_t<float>(rets, covmat_sqrt, gamma_sqrt, alpha, max_weight, B, (int)N, w, state, status, workspace, workspace_bytes, st, batch_total, batch_offset, saved);
}

int ddw_markowitz_fwd_chunk(const void *rets, const void *covmat_sqrt, const void *gamma_sqrt, const void *alpha,
                            double max_weight, int64_t B, int64_t N, int dtype, int64_t batch_total, int64_t batch_offset,
                            void *w, int8_t *state, int32_t *status, void *workspace, size_t workspace_bytes,
                            void *stream) {
    return markowitz_fwd_impl(rets, covmat_sqrt, gamma_sqrt, alpha, max_weight, B, N, dtype, batch_total, batch_offset, w, state,
                              status, workspace, workspace_bytes, nullptr, 0, stream);
}

int ddw_markowitz_fwd_saved(const void *rets, const void *covmat_sqrt, const void *gamma_sqrt, const void *alpha,
                            double max_weight, int64_t B, int64_t N, int dtype, void *w, int8_t *state, int32_t *status,
                            void *workspace, size_t workspace_bytes, void *saved, size_t saved_bytes, void *stream) {
    if (!saved) { g_err[0] = 0; set_error("ddw_markowitz_fwd_saved: null saved buffer"); return DDW_ERR_ARG; }
    return markowitz_fwd_impl(rets, covmat_sqrt, gamma_sqrt, alpha, max_weight, B, N, dtype, B, 0, w, state, status, workspace,
                              workspace_bytes, saved, saved_bytes, stream);
}

int ddw_markowitz_fwd(const void *rets, const void *covmat_sqrt, const void *gamma_sqrt, const void *alpha,
                      double max_weight, int64_t B, int64_t N, int dtype, void *w, int8_t *state, int32_t *status,
                      void *workspace, size_t workspace_bytes, void *stream) {
    return ddw_markowitz_fwd_chunk(rets, covmat_sqrt, gamma_sqrt, alpha, max_weight, B, N, dtype, B, 0, w, state, status,
                                   workspace, workspace_bytes, stream);
}

static int markowitz_bwd_impl(const void *grad_w, const void *w, const int8_t *state, const void *covmat_sqrt,
                              const void *gamma_sqrt, const void *alpha, double max_weight, int64_t B, int64_t N, int dtype,
                              int64_t batch_total, int64_t batch_offset, void *d_rets, void *d_covmat_sqrt,
                              void *d_gamma_sqrt, int accumulate_gamma, void *d_alpha, void *workspace,
                              size_t workspace_bytes, const void *saved, size_t saved_bytes, void *stream) {
    (void)max_weight;
    g_err[0] = 0;
    int rc = check_dims("ddw_markowitz_bwd", B, N, 1024, dtype);
    if (rc) return rc;
    if (B == 0) return 0;
    if (batch_offset < 0 || batch_offset + B > batch_total || batch_total >= (1LL << 31)) { set_error("ddw_markowitz_bwd: chunk [%lld, %lld) outside the batch of %lld", (long long)batch_offset, (long long)(batch_offset + B), (long long)batch_total); return DDW_ERR_ARG; }
    if (!grad_w || !w || !state || !covmat_sqrt || !gamma_sqrt || !alpha || !d_rets || !d_covmat_sqrt || !d_alpha) { set_error("ddw_markowitz_bwd: null pointer"); return DDW_ERR_ARG; }
    DDW_CHECK_ALIGNED("ddw_markowitz_bwd", grad_w, w, covmat_sqrt, gamma_sqrt, alpha, d_rets, d_covmat_sqrt, d_gamma_sqrt, d_alpha, workspace, saved);
    if (saved && saved_bytes < ddw_markowitz_saved_bytes(B, N, dtype)) { set_error("ddw_markowitz_bwd_saved: saved buffer too small"); return DDW_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == DDW_F64
               ? markowitz_bwd_t<double>(grad_w, w, state, covmat_sqrt, gamma_sqrt, alpha, B, (int)N, d_rets, d_covmat_sqrt, d_gamma_sqrt, d_alpha, workspace, workspace_bytes, st, batch_total, batch_offset, accumulate_gamma, saved, nullptr)
               : markowitz_bwd_t<float>(grad_w, w, state, covmat_sqrt, gamma_sqrt, alpha, B, (int)N, d_rets, d_covmat_sqrt, d_gamma_sqrt, d_alpha, workspace, workspace_bytes, st, batch_total, batch_offset, accumulate_gamma, saved, nullptr);
}

int ddw_markowitz_bwd_chunk(const void *grad_w, const void *w, const int8_t *state, const void *covmat_sqrt,
                            const void *gamma_sqrt, const void *alpha, double max_weight, int64_t B, int64_t N, int dtype,
                            int64_t batch_total, int64_t batch_offset, void *d_rets, void *d_covmat_sqrt,
                            void *d_gamma_sqrt, int accumulate_gamma, void *d_alpha, void *workspace,
                            size_t workspace_bytes, void *stream) {
    return markowitz_bwd_impl(grad_w, w, state, covmat_sqrt, gamma_sqrt, alpha, max_weight, B, N, dtype, batch_total, batch_offset,
                              d_rets, d_covmat_sqrt, d_gamma_sqrt, accumulate_gamma, d_alpha, workspace, workspace_bytes, nullptr, 0,
                              stream);
}

int ddw_markowitz_bwd_saved(const void *grad_w, const void *w, const int8_t *state, const void *covmat_sqrt,
                            const void *gamma_sqrt, const void *alpha, double max_weight, int64_t B, int64_t N, int dtype,
                            void *d_rets, void *d_covmat_sqrt, void *d_gamma_sqrt, void *d_alpha, void *workspace,
                            size_t workspace_bytes, const void *saved, size_t saved_bytes, void *stream) {
    if (!saved) { g_err[0] = 0; set_error("ddw_markowitz_bwd_saved: null saved buffer"); return DDW_ERR_ARG; }
    return markowitz_bwd_impl(grad_w, w, state, covmat_sqrt, gamma_sqrt, alpha, max_weight, B, N, dtype, B, 0, d_rets, d_covmat_sqrt,
                              d_gamma_sqrt, 0, d_alpha, workspace, workspace_bytes, saved, saved_bytes, stream);
}

int ddw_markowitz_bwd(const void *grad_w, const void *w, const int8_t *state, const void *covmat_sqrt,
                      const void *gamma_sqrt, const void *alpha, double max_weight, int64_t B, int64_t N, int dtype,
                      void *d_rets, void *d_covmat_sqrt, void *d_gamma_sqrt, void *d_alpha, void *workspace,
                      size_t workspace_bytes, void *stream) {
    return ddw_markowitz_bwd_chunk(grad_w, w, state, covmat_sqrt, gamma_sqrt, alpha, max_weight, B, N, dtype, B, 0, d_rets,
                                   d_covmat_sqrt, d_gamma_sqrt, 0, d_alpha, workspace, workspace_bytes, stream);
}

static int row_call(const char *fn, int which, const void *x, const void *temperature, double tscalar, double u,
                    int64_t B, int64_t N, int dtype, const void *grad_w, const void *w, void *out, void *d_x,
                    void *d_t, void *stream) {
    g_err[0] = 0;
    int rc = check_dims(fn, B, N, 2048, dtype);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!x || (!temperature && !(tscalar != 0.0))) { set_error("%s: null input or zero temperature", fn); return DDW_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == DDW_F64 ? row_op_t<double>(which, x, temperature, tscalar, u, B, (int)N, grad_w, w, out, d_x, d_t, st)
                            : row_op_t<float>(which, x, temperature, tscalar, u, B, (int)N, grad_w, w, out, d_x, d_t, st);
}

int ddw_covariance_markowitz_fwd(const void *x, const void *coef, int strategy, int64_t L, const void *rets,
                                 const void *gamma_sqrt, const void *alpha, double max_weight, int64_t B, int64_t N, int dtype,
                                 void *w, int8_t *state, int32_t *status, void *workspace, size_t workspace_bytes,
                                 void *saved, size_t saved_bytes, void *covmat_scratch, void *stream) {
    g_err[0] = 0;
    int rc = check_dims("ddw_covariance_markowitz_fwd", B, N, 1024, dtype);
    if (rc) return rc;
    if (B == 0) return 0;
    if (L < 2 || L >= (1LL << 24)) { set_error("ddw_covariance_markowitz_fwd: need 2 <= dim < 2^24 observations, got %lld", (long long)L); return DDW_ERR_ARG; }
    if (strategy < DDW_SHRINK_NONE || strategy > DDW_SHRINK_SCALED_IDENTITY) { set_error("ddw_covariance_markowitz_fwd: bad strategy %d", strategy); return DDW_ERR_ARG; }
    if (!x || !rets || !gamma_sqrt || !alpha || !w || !state || !status || (strategy != DDW_SHRINK_NONE && !coef)) { set_error("ddw_covariance_markowitz_fwd: null pointer"); return DDW_ERR_ARG; }
    DDW_CHECK_ALIGNED("ddw_covariance_markowitz_fwd", x, rets, gamma_sqrt, alpha, w, workspace, saved, covmat_scratch);
    if (saved && saved_bytes < ddw_markowitz_saved_bytes(B, N, dtype)) { set_error("ddw_covariance_markowitz_fwd: saved buffer too small"); return DDW_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == DDW_F64
               ? markowitz_fused_fwd_t<double>(x, coef, strategy, (int)L, rets, gamma_sqrt, alpha, max_weight, B, (int)N, w, state, status, workspace, workspace_bytes, st, saved, covmat_scratch)
               : markowitz_fused_fwd_t<float>(x, coef, strategy, (int)L, rets, gamma_sqrt, alpha, max_weight, B, (int)N, w, state, status, workspace, workspace_bytes, st, saved, covmat_scratch);
}

size_t ddw_risk_budgeting_workspace_bytes(int64_t B, int64_t N, int dtype) {
    if (B <= 0 || N <= 0 || N > 1024) return 0;
    return risk_budgeting_workspace_bytes(B, (int)N, dtype == DDW_F64 ? 8 : 4);
}

int ddw_risk_budgeting_fwd(const void *covmat_sqrt, const void *b, double max_weight, int64_t B, int64_t N, int dtype,
                           void *w, int8_t *state, int32_t *status, void *workspace, size_t workspace_bytes, void *stream) {
    g_err[0] = 0;
    int rc = check_dims("ddw_risk_budgeting_fwd", B, N, 1024, dtype);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!covmat_sqrt || !b || !w || !state || !status) { set_error("ddw_risk_budgeting_fwd: null pointer"); return DDW_ERR_ARG; }
    DDW_CHECK_ALIGNED("ddw_risk_budgeting_fwd", covmat_sqrt, b, w, workspace);
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == DDW_F64
               ? risk_budgeting_t<double>(false, covmat_sqrt, b, nullptr, nullptr, nullptr, max_weight, B, (int)N, w, state, status, nullptr, nullptr, workspace, workspace_bytes, st)
               : risk_budgeting_t<float>(false, covmat_sqrt, b, nullptr, nullptr, nullptr, max_weight, B, (int)N, w, state, status, nullptr, nullptr, workspace, workspace_bytes, st);
}

int ddw_risk_budgeting_bwd(const void *grad_w, const void *w, const int8_t *state, const void *covmat_sqrt, const void *b,
                           double max_weight, int64_t B, int64_t N, int dtype, void *d_covmat_sqrt, void *d_b,
                           void *workspace, size_t workspace_bytes, void *stream) {
    g_err[0] = 0;
    int rc = check_dims("ddw_risk_budgeting_bwd", B, N, 1024, dtype);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!grad_w || !w || !state || !covmat_sqrt || !b || !d_covmat_sqrt || !d_b) { set_error("ddw_risk_budgeting_bwd: null pointer"); return DDW_ERR_ARG; }
    DDW_CHECK_ALIGNED("ddw_risk_budgeting_bwd", grad_w, w, covmat_sqrt, b, d_covmat_sqrt, d_b, workspace);
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == DDW_F64
               ? risk_budgeting_t<double>(true, covmat_sqrt, b, grad_w, w, state, max_weight, B, (int)N, nullptr, nullptr, nullptr, d_covmat_sqrt, d_b, workspace, workspace_bytes, st)
               : risk_budgeting_t<float>(true, covmat_sqrt, b, grad_w, w, state, max_weight, B, (int)N, nullptr, nullptr, nullptr, d_covmat_sqrt, d_b, workspace, workspace_bytes, st);
}

int ddw_capped_simplex_fwd(const void *x, const void *temperature, double temperature_scalar, double max_weight,
                           int64_t B, int64_t N, int dtype, void *w, void *stream) {
    if (!w) { set_error("ddw_capped_simplex_fwd: null output"); return DDW_ERR_ARG; }
    return row_call("ddw_capped_simplex_fwd", 0, x, temperature, temperature_scalar, max_weight, B, N, dtype, nullptr, nullptr, w, nullptr, nullptr, stream);
}

int ddw_capped_simplex_bwd(const void *grad_w, const void *w, const void *x, const void *temperature,
                           double temperature_scalar, double max_weight, int64_t B, int64_t N, int dtype, void *d_x,
                           void *d_temperature, void *stream) {
    if (!grad_w || !w || !d_x) { set_error("ddw_capped_simplex_bwd: null pointer"); return DDW_ERR_ARG; }
    return row_call("ddw_capped_simplex_bwd", 1, x, temperature, temperature_scalar, max_weight, B, N, dtype, grad_w, w, nullptr, d_x, d_temperature, stream);
}

int ddw_capped_argmax_fwd(const void *x, double max_weight, int64_t B, int64_t N, int dtype, void *w, void *stream) {
    if (!w) { set_error("ddw_capped_argmax_fwd: null output"); return DDW_ERR_ARG; }
    if (!(max_weight > 0.0) || (double)N * max_weight < 1.0 - 1e-12) {
        set_error("ddw_capped_argmax_fwd: n_assets * max_weight < 1, no fully invested portfolio exists");
        return DDW_ERR_ARG;
    }
    return row_call("ddw_capped_argmax_fwd", 4, x, nullptr, 1.0, max_weight, B, N, dtype, nullptr, nullptr, w, nullptr, nullptr, stream);
}

int ddw_capped_softmax_fwd(const void *x, const void *temperature, double temperature_scalar, double max_weight,
                           int64_t B, int64_t N, int dtype, void *w, void *stream) {
    if (!w) { set_error("ddw_capped_softmax_fwd: null output"); return DDW_ERR_ARG; }
    return row_call("ddw_capped_softmax_fwd", 2, x, temperature, temperature_scalar, max_weight, B, N, dtype, nullptr, nullptr, w, nullptr, nullptr, stream);
}

int ddw_capped_softmax_bwd(const void *grad_w, const void *w, const void *x, const void *temperature,
                           double temperature_scalar, double max_weight, int64_t B, int64_t N, int dtype, void *d_x,
                           void *d_temperature, void *stream) {
    if (!grad_w || !w || !d_x) { set_error("ddw_capped_softmax_bwd: null pointer"); return DDW_ERR_ARG; }
    return row_call("ddw_capped_softmax_bwd", 3, x, temperature, temperature_scalar, max_weight, B, N, dtype, grad_w, w, nullptr, d_x, d_temperature, stream);
}

size_t ddw_covariance_workspace_bytes(int64_t B, int64_t L, int64_t N, int dtype, int do_sqrt) {
    if (B <= 0 || N <= 0 || L <= 0 || N > 1024) return 0;
    return covariance_workspace_bytes(B, (int)L, (int)N, dtype == DDW_F64 ? 8 : 4, do_sqrt);
}

int ddw_covariance_fwd(const void *x, const void *coef, int strategy, int do_sqrt, int64_t B, int64_t L, int64_t N,
                       int dtype, void *out, void *eigvec, void *eigval, int32_t *sweeps, void *workspace,
                       size_t workspace_bytes, void *stream) {
    g_err[0] = 0;
    int rc = check_dims("ddw_covariance_fwd", B, N, 1024, dtype);
    if (rc) return rc;
    if (B == 0) return 0;
    if (L < 2 || L >= (1LL << 24)) { set_error("ddw_covariance_fwd: need 2 <= dim < 2^24 observations, got %lld", (long long)L); return DDW_ERR_ARG; }
    if (strategy < DDW_SHRINK_NONE || strategy > DDW_SHRINK_SCALED_IDENTITY) { set_error("ddw_covariance_fwd: bad strategy %d", strategy); return DDW_ERR_ARG; }
    if (!x || !out || (strategy != DDW_SHRINK_NONE && !coef)) { set_error("ddw_covariance_fwd: null pointer"); return DDW_ERR_ARG; }
    DDW_CHECK_ALIGNED("ddw_covariance_fwd", x, out, eigvec, eigval, workspace);
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == DDW_F64
               ? covariance_fwd_t<double>(x, coef, strategy, do_sqrt, B, (int)L, (int)N, out, eigvec, eigval, sweeps, workspace, workspace_bytes, st)
               : covariance_fwd_t<float>(x, coef, strategy, do_sqrt, B, (int)L, (int)N, out, eigvec, eigval, sweeps, workspace, workspace_bytes, st);
}

int ddw_covariance_bwd(const void *grad_out, const void *x, const void *coef, int strategy, int do_sqrt,
                       const void *eigvec, const void *eigval, int64_t B, int64_t L, int64_t N, int dtype, void *d_x,
                       void *d_coef, void *workspace, size_t workspace_bytes, void *stream) {
    g_err[0] = 0;
    int rc = check_dims("ddw_covariance_bwd", B, N, 1024, dtype);
    if (rc) return rc;
    if (B == 0) return 0;
    if (L < 2 || L >= (1LL << 24)) { set_error("ddw_covariance_bwd: need 2 <= dim < 2^24 observations, got %lld", (long long)L); return DDW_ERR_ARG; }
    if (strategy < DDW_SHRINK_NONE || strategy > DDW_SHRINK_SCALED_IDENTITY) { set_error("ddw_covariance_bwd: bad strategy %d", strategy); return DDW_ERR_ARG; }
    if (!grad_out || !x || !d_x || (strategy != DDW_SHRINK_NONE && !coef) || (do_sqrt && (!eigvec || !eigval))) { set_error("ddw_covariance_bwd: null pointer"); return DDW_ERR_ARG; }
    DDW_CHECK_ALIGNED("ddw_covariance_bwd", grad_out, x, eigvec, eigval, d_x, workspace);
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == DDW_F64
               ? covariance_bwd_t<double>(grad_out, x, coef, strategy, do_sqrt, eigvec, eigval, B, (int)L, (int)N, d_x, d_coef, workspace, workspace_bytes, st)
               : covariance_bwd_t<float>(grad_out, x, coef, strategy, do_sqrt, eigvec, eigval, B, (int)L, (int)N, d_x, d_coef, workspace, workspace_bytes, st);
}

/* psd sqrt of given symmetric matrices: the covariance kernels in direct mode (L = 0, no shrinkage) */
int ddw_psd_sqrt_fwd(const void *m, int64_t B, int64_t N, int dtype, void *out, void *eigvec, void *eigval,
                     void *workspace, size_t workspace_bytes, void *stream) {
    g_err[0] = 0;
    int rc = check_dims("ddw_psd_sqrt_fwd", B, N, 1024, dtype);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!m || !out) { set_error("ddw_psd_sqrt_fwd: null pointer"); return DDW_ERR_ARG; }
    DDW_CHECK_ALIGNED("ddw_psd_sqrt_fwd", m, out, eigvec, eigval, workspace);
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == DDW_F64
               ? covariance_fwd_t<double>(m, nullptr, DDW_SHRINK_NONE, 1, B, 0, (int)N, out, eigvec, eigval, nullptr, workspace, workspace_bytes, st)
               : covariance_fwd_t<float>(m, nullptr, DDW_SHRINK_NONE, 1, B, 0, (int)N, out, eigvec, eigval, nullptr, workspace, workspace_bytes, st);
}

int ddw_psd_sqrt_bwd(const void *grad_out, const void *eigvec, const void *eigval, int64_t B, int64_t N, int dtype,
                     void *d_m, void *workspace, size_t workspace_bytes, void *stream) {
    g_err[0] = 0;
    int rc = check_dims("ddw_psd_sqrt_bwd", B, N, 1024, dtype);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!grad_out || !eigvec || !eigval || !d_m) { set_error("ddw_psd_sqrt_bwd: null pointer"); return DDW_ERR_ARG; }
    DDW_CHECK_ALIGNED("ddw_psd_sqrt_bwd", grad_out, eigvec, eigval, d_m, workspace);
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == DDW_F64
               ? covariance_bwd_t<double>(grad_out, grad_out, nullptr, DDW_SHRINK_NONE, 1, eigvec, eigval, B, 0, (int)N, d_m, nullptr, workspace, workspace_bytes, st)
               : covariance_bwd_t<float>(grad_out, grad_out, nullptr, DDW_SHRINK_NONE, 1, eigvec, eigval, B, 0, (int)N, d_m, nullptr, workspace, workspace_bytes, st);
}

}  // extern "C"
