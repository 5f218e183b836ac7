// host_pipeline.cu -- HOST-buffer entry point of NumericalMarkowitz (include/ddw.h, "HOST-buffer entry point").
//
// The reference layer is called with CPU tensors and returns CPU tensors (deepdow/layers/allocate.py:240-273,
// solved on host threads behind CvxpyLayer, :273).  A caller that keeps its data on the host therefore needs
// host -> device -> host around the solver kernels.  Done naively that is three serial phases (copy in, solve,
// copy out) and PCIe, not the GPU, sets the pace: at B = 4096, N = 100 the step moves 167 MB each way.
// Here the (B,N,N) tensors -- 98% of the bytes -- are cut into chunks that flow through three streams
//     s_in : H2D of chunk i+1      s_run[i % 2] : forward + backward of chunk i      s_out : D2H of chunk i-1
// so that both PCIe directions and the kernels run at the same time; the small per-problem vectors travel once
// per batch.  Chunk slots live in caller-owned device scratch and are recycled round-robin, guarded by events.
//
// A step enqueues ~15 operations per chunk, and at 16-32 chunks the host cannot issue them as fast as the GPU
// retires them (measured: 70-100 us of enqueue time per chunk).  The whole step -- copies, kernels and the
// cross-stream dependencies -- is therefore captured into a CUDA graph the second time it is requested with the
// same arguments and replayed with a single launch from then on.  Capture needs page-locked host buffers;
// pageable ones take the eager path.
#include <stdlib.h>

#include "common.cuh"

namespace ddw {

size_t markowitz_workspace_bytes(long long B, int N, int tsize);
template <typename T>
int markowitz_fwd_t(const void *, const void *, const void *, const void *, double, long long, int, void *, int8_t *,
                    int32_t *, void *, size_t, cudaStream_t, long long, long long, void * = nullptr);
template <typename T>
int markowitz_bwd_t(const void *, const void *, const int8_t *, const void *, const void *, const void *, long long,
                    int, void *, void *, void *, void *, void *, size_t, cudaStream_t, long long, long long, int,
                    const void * = nullptr, void * = nullptr);

constexpr int kMaxSlots = 8, kMaxRun = 4;
// chunk slots in the device scratch / solver streams (consecutive chunks are solved concurrently: a chunk alone cannot
// fill the GPU).  Environment overrides are for tuning runs (tools/time_e2e.py).
static int env_int(const char *name, int dflt, int lo, int hi) {
    const char *v = getenv(name);
    if (!v) return dflt;
    const int x = atoi(v);
    return x < lo ? lo : (x > hi ? hi : x);
}
static const int kSlots = env_int("DDW_HP_SLOTS", 4, 2, kMaxSlots);
static const int kRun = env_int("DDW_HP_RUN", 2, 1, kMaxRun);

static size_t up256(size_t v) { return (v + 255) / 256 * 256; }

// device scratch: whole-batch vectors, then kSlots chunk slots of (covmat_sqrt, d_covmat_sqrt, workspace)
struct ScratchLayout {
    size_t gamma, d_gamma, rets, alpha, grad_w, w, state, status, d_rets, d_alpha, vecs;   // whole batch
    size_t slots, slot_bytes, C, d_C, ws, ws_bytes;                                  // per slot
    size_t total;
};

static ScratchLayout scratch_layout(long long B, long long chunk, int N, int tsize) {
    ScratchLayout l;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += up256(bytes); return at; };
    const size_t b = (size_t)B, c = (size_t)chunk, n = (size_t)N;
    l.gamma = take(b * tsize);
    l.d_gamma = take(b * tsize);
    l.rets = take(b * n * tsize);
    l.alpha = take(b * tsize);
    l.grad_w = take(b * n * tsize);
    l.w = take(b * n * tsize);
    l.state = take(b * n);
    l.status = take(b * sizeof(int32_t));
    l.d_rets = take(b * n * tsize);
    l.d_alpha = take(b * tsize);
    l.vecs = take(b * 2 * n * tsize);      // (A d, A w): the factored form of d_covmat_sqrt
    l.slots = o;
    o = 0;
    l.C = take(c * n * n * tsize);
    l.d_C = take(c * n * n * tsize);
    l.ws_bytes = markowitz_workspace_bytes(chunk, N, tsize);
    l.ws = take(l.ws_bytes);
    l.slot_bytes = o;
    l.total = l.slots + (size_t)kSlots * l.slot_bytes;
    return l;
}

struct StepArgs {
    const void *h_rets, *h_covmat_sqrt, *h_gamma_sqrt, *h_alpha, *h_grad_w;
    double max_weight;
    int64_t B, N, chunk;
    int dtype;
    void *h_w;
    int32_t *h_status;
    void *h_d_rets, *h_d_covmat_sqrt, *h_d_gamma_sqrt, *h_d_alpha, *dev_scratch;
    size_t dev_scratch_bytes;
    void *h_factors;     // non-null: return (A d, A w) per problem instead of d_covmat_sqrt (ddw_markowitz_host_step_factored)
    bool operator==(const StepArgs &o) const {
        return h_rets == o.h_rets && h_covmat_sqrt == o.h_covmat_sqrt && h_gamma_sqrt == o.h_gamma_sqrt &&
               h_alpha == o.h_alpha && h_grad_w == o.h_grad_w && max_weight == o.max_weight && B == o.B && N == o.N &&
               chunk == o.chunk && dtype == o.dtype && h_w == o.h_w && h_status == o.h_status && h_d_rets == o.h_d_rets &&
               h_d_covmat_sqrt == o.h_d_covmat_sqrt && h_d_gamma_sqrt == o.h_d_gamma_sqrt && h_d_alpha == o.h_d_alpha &&
               dev_scratch == o.dev_scratch && dev_scratch_bytes == o.dev_scratch_bytes && h_factors == o.h_factors;
    }
};

}  // namespace ddw

using namespace ddw;

struct ddw_host_pipeline {
    cudaStream_t s_in, s_run[kMaxRun], s_out, s_origin;
    cudaEvent_t start, in_done[kMaxSlots], run_done[kMaxSlots], out_done[kMaxSlots], head_in, head_run, tail_in, tail_run[kMaxRun], tail_out;
    // graph cache: the last argument set seen, and its instantiated graph once it has been seen twice
    StepArgs last;
    bool have_last;
    cudaGraphExec_t exec;
    int graph_launches;
};

namespace ddw {

static bool page_locked(const void *p) {
    if (!p) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// Enqueue one step: everything forks from `origin` and joins back into it.
static int enqueue_step(ddw_host_pipeline *pipe, const StepArgs &a, cudaStream_t origin) {
    const int ts = a.dtype == DDW_F64 ? 8 : 4, n = (int)a.N;
    const int64_t B = a.B, chunk = a.chunk;
    const ScratchLayout L = scratch_layout(B, chunk, n, ts);
    unsigned char *base = reinterpret_cast<unsigned char *>(a.dev_scratch);
    const bool bwd = a.h_grad_w != nullptr;
    const bool want_dgamma = bwd && a.h_d_gamma_sqrt != nullptr;
    const size_t row = (size_t)n * ts, mat = (size_t)n * n * ts;

    cudaEventRecord(pipe->start, origin);
    cudaStreamWaitEvent(pipe->s_in, pipe->start, 0);
    for (int r = 0; r < kRun; ++r) cudaStreamWaitEvent(pipe->s_run[r], pipe->start, 0);
    cudaStreamWaitEvent(pipe->s_out, pipe->start, 0);

    // per-problem vectors of the whole batch: 2% of the bytes, one copy each
    cudaMemcpyAsync(base + L.gamma, a.h_gamma_sqrt, (size_t)B * ts, cudaMemcpyHostToDevice, pipe->s_in);
    cudaMemcpyAsync(base + L.rets, a.h_rets, (size_t)B * row, cudaMemcpyHostToDevice, pipe->s_in);
    cudaMemcpyAsync(base + L.alpha, a.h_alpha, (size_t)B * ts, cudaMemcpyHostToDevice, pipe->s_in);
    if (bwd) cudaMemcpyAsync(base + L.grad_w, a.h_grad_w, (size_t)B * row, cudaMemcpyHostToDevice, pipe->s_in);
    cudaEventRecord(pipe->head_in, pipe->s_in);
    cudaStreamWaitEvent(pipe->s_run[0], pipe->head_in, 0);
    if (want_dgamma) cudaMemsetAsync(base + L.d_gamma, 0, (size_t)B * ts, pipe->s_run[0]);
    cudaEventRecord(pipe->head_run, pipe->s_run[0]);
    for (int r = 1; r < kRun; ++r) cudaStreamWaitEvent(pipe->s_run[r], pipe->head_run, 0);

    // Chunk schedule: full chunks in the middle, a quarter and a half chunk at both ends -- the first kernels start
    // after a quarter chunk has arrived instead of a whole one, and the last copy-out is a quarter chunk long
    // (fill and drain are the only parts of the step that do not overlap).
    static const bool no_ramp = getenv("DDW_HOST_PIPELINE_NO_RAMP") != nullptr;   // diagnostic
    const bool ramp = !no_ramp && B >= 4 * chunk && chunk >= 4;
    const int64_t q4 = chunk / 4, q2 = chunk / 2;
    int rc = 0, it = 0;
    int64_t c = 0;
    for (int64_t b0 = 0; b0 < B; b0 += c, ++it) {
        const int64_t left = B - b0;
        c = left < chunk ? left : chunk;
        if (ramp) {
            if (it == 0) c = q4;
            else if (it == 1) c = q2;
            else if (left <= q4) c = left;
            else if (left <= q2 + q4) c = left - q4;
            else if (left <= chunk + q2 + q4) c = left - (q2 + q4);
        }
        const int sl = it % kSlots;
        cudaStream_t run = pipe->s_run[it % kRun];
        unsigned char *S = base + L.slots + (size_t)sl * L.slot_bytes;
        // the slot is free once its previous tenant has been copied out
        if (it >= kSlots) cudaStreamWaitEvent(pipe->s_in, pipe->out_done[sl], 0);
        cudaMemcpyAsync(S + L.C, reinterpret_cast<const unsigned char *>(a.h_covmat_sqrt) + b0 * mat, c * mat,
                        cudaMemcpyHostToDevice, pipe->s_in);
        cudaEventRecord(pipe->in_done[sl], pipe->s_in);

        cudaStreamWaitEvent(run, pipe->in_done[sl], 0);
        unsigned char *rets = base + L.rets + b0 * row, *alpha = base + L.alpha + b0 * ts, *w = base + L.w + b0 * row;
        unsigned char *grad_w = base + L.grad_w + b0 * row, *d_rets = base + L.d_rets + b0 * row;
        unsigned char *d_alpha = base + L.d_alpha + b0 * ts;
        int8_t *state = reinterpret_cast<int8_t *>(base + L.state) + b0 * n;
        int32_t *status = reinterpret_cast<int32_t *>(base + L.status) + b0;
        if (a.dtype == DDW_F64)
            rc = markowitz_fwd_t<double>(rets, S + L.C, base + L.gamma, alpha, a.max_weight, c, n, w, state, status, S + L.ws, L.ws_bytes, run, B, b0);
        else
            rc = markowitz_fwd_t<float>(rets, S + L.C, base + L.gamma, alpha, a.max_weight, c, n, w, state, status, S + L.ws, L.ws_bytes, run, B, b0);
        if (!rc && bwd) {
            void *dg = want_dgamma ? base + L.d_gamma : nullptr;
            void *vo = a.h_factors ? base + L.vecs + b0 * 2 * row : nullptr;
            if (a.dtype == DDW_F64)
                rc = markowitz_bwd_t<double>(grad_w, w, state, S + L.C, base + L.gamma, alpha, c, n, d_rets, S + L.d_C, dg, d_alpha, S + L.ws, L.ws_bytes, run, B, b0, 1, nullptr, vo);
            else
                rc = markowitz_bwd_t<float>(grad_w, w, state, S + L.C, base + L.gamma, alpha, c, n, d_rets, S + L.d_C, dg, d_alpha, S + L.ws, L.ws_bytes, run, B, b0, 1, nullptr, vo);
        }
        cudaEventRecord(pipe->run_done[sl], run);
        if (rc) break;

        cudaStreamWaitEvent(pipe->s_out, pipe->run_done[sl], 0);
        if (bwd && !a.h_factors)
            cudaMemcpyAsync(reinterpret_cast<unsigned char *>(a.h_d_covmat_sqrt) + b0 * mat, S + L.d_C, c * mat,
                            cudaMemcpyDeviceToHost, pipe->s_out);
        cudaEventRecord(pipe->out_done[sl], pipe->s_out);
    }
    if (!rc) {
        // s_out waits for the tail of every solver stream, so every chunk's results are final
        for (int r = 0; r < kRun; ++r) { cudaEventRecord(pipe->tail_run[r], pipe->s_run[r]); cudaStreamWaitEvent(pipe->s_out, pipe->tail_run[r], 0); }
        cudaMemcpyAsync(a.h_w, base + L.w, (size_t)B * row, cudaMemcpyDeviceToHost, pipe->s_out);
        cudaMemcpyAsync(a.h_status, base + L.status, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, pipe->s_out);
        if (bwd) {
            cudaMemcpyAsync(a.h_d_rets, base + L.d_rets, (size_t)B * row, cudaMemcpyDeviceToHost, pipe->s_out);
            cudaMemcpyAsync(a.h_d_alpha, base + L.d_alpha, (size_t)B * ts, cudaMemcpyDeviceToHost, pipe->s_out);
            if (a.h_factors) cudaMemcpyAsync(a.h_factors, base + L.vecs, (size_t)B * 2 * row, cudaMemcpyDeviceToHost, pipe->s_out);
        }
        if (want_dgamma) cudaMemcpyAsync(a.h_d_gamma_sqrt, base + L.d_gamma, (size_t)B * ts, cudaMemcpyDeviceToHost, pipe->s_out);
    }
    // `origin` resumes when all three streams have drained
    cudaEventRecord(pipe->tail_in, pipe->s_in);
    for (int r = 0; r < kRun; ++r) cudaEventRecord(pipe->tail_run[r], pipe->s_run[r]);
    cudaEventRecord(pipe->tail_out, pipe->s_out);
    cudaStreamWaitEvent(origin, pipe->tail_in, 0);
    for (int r = 0; r < kRun; ++r) cudaStreamWaitEvent(origin, pipe->tail_run[r], 0);
    cudaStreamWaitEvent(origin, pipe->tail_out, 0);
    return rc;
}

}  // namespace ddw

extern "C" {

int ddw_host_pipeline_create(ddw_host_pipeline **out) {
    if (!out) { set_error("ddw_host_pipeline_create: null pointer"); return DDW_ERR_ARG; }
    ddw_host_pipeline *p = new ddw_host_pipeline();
    p->have_last = false; p->exec = nullptr; p->graph_launches = 0;
    bool ok = cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&p->s_origin, cudaStreamNonBlocking) == cudaSuccess;
    for (int r = 0; r < kRun; ++r) ok = ok && cudaStreamCreateWithFlags(&p->s_run[r], cudaStreamNonBlocking) == cudaSuccess;
    auto ev = [&](cudaEvent_t *e) { ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess; };
    ev(&p->start); ev(&p->head_in); ev(&p->head_run); ev(&p->tail_in); ev(&p->tail_out);
    for (int r = 0; r < kRun; ++r) ev(&p->tail_run[r]);
    for (int i = 0; i < kSlots; ++i) { ev(&p->in_done[i]); ev(&p->run_done[i]); ev(&p->out_done[i]); }
    if (!ok) {
        check_launch("ddw_host_pipeline_create");
        delete p;
        return DDW_ERR_CUDA;
    }
    *out = p;
    return 0;
}

void ddw_host_pipeline_destroy(ddw_host_pipeline *p) {
    if (!p) return;
    cudaStreamSynchronize(p->s_in); cudaStreamSynchronize(p->s_out);
    for (int r = 0; r < kRun; ++r) cudaStreamSynchronize(p->s_run[r]);
    cudaStreamSynchronize(p->s_origin);
    if (p->exec) cudaGraphExecDestroy(p->exec);
    cudaEventDestroy(p->start); cudaEventDestroy(p->head_in); cudaEventDestroy(p->head_run);
    cudaEventDestroy(p->tail_in); cudaEventDestroy(p->tail_out);
    for (int r = 0; r < kRun; ++r) { cudaEventDestroy(p->tail_run[r]); cudaStreamDestroy(p->s_run[r]); }
    for (int i = 0; i < kSlots; ++i) { cudaEventDestroy(p->in_done[i]); cudaEventDestroy(p->run_done[i]); cudaEventDestroy(p->out_done[i]); }
    cudaStreamDestroy(p->s_in); cudaStreamDestroy(p->s_out); cudaStreamDestroy(p->s_origin);
    delete p;
}

/* number of steps served by replaying the cached CUDA graph (diagnostic) */
int ddw_host_pipeline_graph_launches(const ddw_host_pipeline *p) { return p ? p->graph_launches : 0; }

size_t ddw_markowitz_host_scratch_bytes(int64_t B, int64_t N, int dtype, int64_t chunk) {
    if (B <= 0 || N <= 0 || N > 1024 || chunk <= 0) return 0;
    if (chunk > B) chunk = B;
    return scratch_layout(B, chunk, (int)N, dtype == DDW_F64 ? 8 : 4).total;
}

static int host_step_impl(ddw_host_pipeline *pipe, const void *h_rets, const void *h_covmat_sqrt,
                          const void *h_gamma_sqrt, const void *h_alpha, const void *h_grad_w, double max_weight,
                          int64_t B, int64_t N, int dtype, int64_t chunk, void *h_w, int32_t *h_status, void *h_d_rets,
                          void *h_d_covmat_sqrt, void *h_factors, void *h_d_gamma_sqrt, void *h_d_alpha, void *dev_scratch,
                          size_t dev_scratch_bytes, void *stream) {
    const char *fn = "ddw_markowitz_host_step";
    if (dtype != DDW_F32 && dtype != DDW_F64) { set_error("%s: dtype must be DDW_F32 or DDW_F64", fn); return DDW_ERR_ARG; }
    if (B < 0 || B >= (1LL << 31) || chunk <= 0) { set_error("%s: bad batch size / chunk", fn); return DDW_ERR_ARG; }
    if (N < 1 || N > 1024) { set_error("%s: n_assets=%lld outside [1, 1024]", fn, (long long)N); return DDW_ERR_UNSUPPORTED; }
    if (B == 0) return 0;
    const bool bwd = h_grad_w != nullptr;
    if (!pipe || !h_rets || !h_covmat_sqrt || !h_gamma_sqrt || !h_alpha || !h_w || !h_status ||
        (bwd && (!h_d_rets || (!h_d_covmat_sqrt && !h_factors) || !h_d_alpha))) { set_error("%s: null pointer", fn); return DDW_ERR_ARG; }
    if (chunk > B) chunk = B;
    if (!dev_scratch || dev_scratch_bytes < ddw_markowitz_host_scratch_bytes(B, N, dtype, chunk)) { set_error("%s: device scratch too small", fn); return DDW_ERR_WORKSPACE; }
    if (reinterpret_cast<uintptr_t>(dev_scratch) & 255u) { set_error("%s: device scratch must be 256-byte aligned", fn); return DDW_ERR_ARG; }
    if (!bwd) h_d_rets = h_d_covmat_sqrt = h_d_gamma_sqrt = h_d_alpha = h_factors = nullptr;

    const StepArgs a = {h_rets, h_covmat_sqrt, h_gamma_sqrt, h_alpha, h_grad_w, max_weight, B, N, chunk, dtype, h_w, h_status,
                        h_d_rets, h_d_covmat_sqrt, h_d_gamma_sqrt, h_d_alpha, dev_scratch, dev_scratch_bytes, h_factors};
    cudaStream_t user = (cudaStream_t)stream;
    static const bool no_graph = getenv("DDW_HOST_PIPELINE_NO_GRAPH") != nullptr;   // diagnostic: always enqueue eagerly
    const bool same = !no_graph && pipe->have_last && pipe->last == a;
    if (same && pipe->exec) {
        ++pipe->graph_launches;
        cudaGraphLaunch(pipe->exec, user);
        return check_launch(fn);
    }
    if (pipe->exec) { cudaGraphExecDestroy(pipe->exec); pipe->exec = nullptr; }
    if (same && page_locked(h_rets) && page_locked(h_covmat_sqrt) && page_locked(h_gamma_sqrt) && page_locked(h_alpha) &&
        page_locked(h_grad_w) && page_locked(h_w) && page_locked(h_status) && page_locked(h_d_rets) &&
        page_locked(h_d_covmat_sqrt) && page_locked(h_d_gamma_sqrt) && page_locked(h_d_alpha) && page_locked(h_factors)) {
        // second request with these arguments: record the step once, replay it from now on
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(pipe->s_origin, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const int rc = enqueue_step(pipe, a, pipe->s_origin);
            const cudaError_t e = cudaStreamEndCapture(pipe->s_origin, &graph);
            if (!rc && e == cudaSuccess && graph &&
                cudaGraphInstantiate(&pipe->exec, graph, nullptr, nullptr, 0) == cudaSuccess) {
                cudaGraphDestroy(graph);
                ++pipe->graph_launches;
                cudaGraphLaunch(pipe->exec, user);
                return check_launch(fn);
            }
            if (graph) cudaGraphDestroy(graph);
            pipe->exec = nullptr;
        }
        cudaGetLastError();   // capture failed: fall through to the eager path
    }
    pipe->last = a;
    pipe->have_last = true;
    const int rc = enqueue_step(pipe, a, user);
    if (rc) return rc;
    return check_launch(fn);
}

int ddw_markowitz_host_step(ddw_host_pipeline *pipe, const void *h_rets, const void *h_covmat_sqrt,
                            const void *h_gamma_sqrt, const void *h_alpha, const void *h_grad_w, double max_weight,
                            int64_t B, int64_t N, int dtype, int64_t chunk, void *h_w, int32_t *h_status, void *h_d_rets,
                            void *h_d_covmat_sqrt, void *h_d_gamma_sqrt, void *h_d_alpha, void *dev_scratch,
                            size_t dev_scratch_bytes, void *stream) {
    return host_step_impl(pipe, h_rets, h_covmat_sqrt, h_gamma_sqrt, h_alpha, h_grad_w, max_weight, B, N, dtype, chunk, h_w, h_status,
                          h_d_rets, h_d_covmat_sqrt, nullptr, h_d_gamma_sqrt, h_d_alpha, dev_scratch, dev_scratch_bytes, stream);
}

int ddw_markowitz_host_step_factored(ddw_host_pipeline *pipe, const void *h_rets, const void *h_covmat_sqrt,
                                     const void *h_gamma_sqrt, const void *h_alpha, const void *h_grad_w, double max_weight,
                                     int64_t B, int64_t N, int dtype, int64_t chunk, void *h_w, int32_t *h_status, void *h_d_rets,
                                     void *h_factors, void *h_d_gamma_sqrt, void *h_d_alpha, void *dev_scratch,
                                     size_t dev_scratch_bytes, void *stream) {
    if (h_grad_w && !h_factors) { set_error("ddw_markowitz_host_step_factored: null factor buffer"); return DDW_ERR_ARG; }
    return host_step_impl(pipe, h_rets, h_covmat_sqrt, h_gamma_sqrt, h_alpha, h_grad_w, max_weight, B, N, dtype, chunk, h_w, h_status,
                          h_d_rets, nullptr, h_factors, h_d_gamma_sqrt, h_d_alpha, dev_scratch, dev_scratch_bytes, stream);
}

}  // extern "C"
