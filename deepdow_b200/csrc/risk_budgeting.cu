// risk_budgeting.cu -- NumericalRiskBudgeting forward and backward.
//
// Replaces the CvxpyLayer call of the reference's NumericalRiskBudgeting.forward (deepdow/layers/allocate.py:673-691;
// problem stated at :651-671):
//
//   minimise  1/2 ||A w||^2 - b' log(w)   s.t.  1'w = 1, 0 <= w <= u          A = covmat_sqrt, Q = A'A
//
// b > 0 keeps every weight strictly positive, so only the cap can bind.  One CTA per portfolio, Q (lower triangle) and the
// factor of the Newton system share one shared-memory array exactly as in the NumericalMarkowitz dense kernel
// (markowitz_impl.cuh), whose building blocks (Gram, blocked LDL', substitutions) are reused.
// Forward: feasible-point Newton on the free set F = {w < u}: H = Q_FF + diag(b_F / w_F^2), the step keeps 1'w, is cut at
// 99% of the way to w = 0 and exactly at w = u (the blocking asset joins the capped set), backtracking on the objective;
// at a stationary point the cap multipliers mu_i = b_i / u - (Q w)_i - nu decide whether an asset is released.
// Backward: [[H, 1], [1', 0]] [d_F; eta] = [g_F; 0]; d_b = d / w on F; d_covmat_sqrt = -[(A d) w' + (A w) d'].
#include "markowitz_common.cuh"

namespace ddw {

constexpr int kMaxRbIter = 300;

template <typename T> struct RbParams {
    const T *A, *b, *grad_w, *w_in;
    const int8_t *state_in;
    T u;
    long long B;
    int N, LD, AS, CH;
    T *w, *d_A, *d_b;
    int8_t *state;
    int32_t *status;
    int *bad_count;
    T *ws;            // global Q / factor storage when the matrix does not fit shared memory
};

template <typename T> __device__ __forceinline__ T logT(T v);
template <> __device__ __forceinline__ float logT<float>(float v) { return logf(v); }
template <> __device__ __forceinline__ double logT<double>(double v) { return log(v); }

// ordered list of the free assets (warp 0); returns through misc[0]
template <typename T>
__device__ __forceinline__ void rb_free_list(const MkSmem<T> &s, int N) {
    const int lane = threadIdx.x & 31;
    int nf = 0;
    for (int base = 0; base < N; base += 32) {
        const int i = base + lane;
        const int v = i < N ? (int)s.st[i] : 2;
        const unsigned mf = __ballot_sync(kFullMask, v == 0);
        if (v == 0) s.idx[nf + __popc(mf & ((1u << lane) - 1u))] = i;
        nf += __popc(mf);
    }
    if (lane == 0) s.misc[0] = nf;
}

template <typename T, int kThreads, int NREG, bool kGlobal>
__global__ void __launch_bounds__(kThreads) risk_budgeting_fwd_kernel(RbParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.N, LD = p.LD, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;
    const long long b = blockIdx.x;
    const T u = p.u;
    T *w_out = p.w + b * N;
    int8_t *st_out = p.state + b * N;
    if (trivial_cases<T, kThreads>(N, u, w_out, st_out, p.status + b, p.bad_count)) return;
    MkSmem<T> s = carve<T, kGlobal>(smem_raw, N, LD, p.AS, p.CH, p.ws, b);
    const int NP = round_up(N + 2, 32);
    T *Qs = s.Qs, *Lb = s.Qs + 1;
    T *bv = s.r, *wv = s.wv, *Qw = s.g;
    gram_lower<T, kThreads, false>(p.A + b * (long long)N * N, nullptr, 0, 1, N, N, nullptr, nullptr, N, s.As, p.AS, p.CH, Qs, LD, T(0));
    T *dw = s.As, *Qd = s.As + NP;          // the Gram's staging area is free from here on (2 CH AS >= 2 NP elements)
    T dmax = T(0);
    for (int i = tid; i < N; i += kThreads) {
        const T bi = p.b[b * N + i];
        bv[i] = bi > Num<T>::tiny() ? bi : Num<T>::tiny();
        wv[i] = T(1) / (T)N;
        s.st[i] = 0;
        dmax = max(dmax, Qs[i * LD + i]);
    }
    dmax = warp_max(dmax);
    if (lane == 0) s.red[warp] = dmax;
    __syncthreads();
    dmax = T(0);
    for (int i = 0; i < kWarps; ++i) dmax = max(dmax, s.red[i]);
    __syncthreads();
    const T dmin = T(4) * Num<T>::eps() * dmax + Num<T>::tiny();
    const T eps = Num<T>::eps();
    int lifted = 0, converged = 0, nonfinite = 0, iters = 0;
    T nu = T(0);
    for (int it = 0; it < kMaxRbIter; ++it) {
        iters = it + 1;
        if (warp == 0) rb_free_list(s, N);
        // Q w for every asset
        for (int i = tid; i < N; i += kThreads) {
            T acc = T(0);
            for (int j = 0; j < N; ++j) acc = fma(sym_at(Qs, LD, i, j), wv[j], acc);
            Qw[i] = acc;
        }
        __syncthreads();
        const int nF = s.misc[0];
        // H = Q_FF + diag(b / w^2) (column-major lower triangle) and the right-hand sides g0 = (Q w - b / w)_F, 1
        for (int cc = warp; cc < nF; cc += kWarps) {
            const int jc = s.idx[cc];
            for (int a = cc + lane; a < nF; a += 32) {
                T v = Qs[s.idx[a] * LD + jc];
                if (a == cc) { const T wi = wv[jc]; v += bv[jc] / (wi * wi); }
                Lb[cc * LD + a] = v;
            }
        }
        for (int a = tid; a < nF; a += kThreads) {
            const int i = s.idx[a];
            Lb[a * LD + nF] = Qw[i] - bv[i] / wv[i];
            Lb[a * LD + nF + 1] = T(1);
        }
        __syncthreads();
        lifted |= factor_ldl<T, kThreads>(Lb, LD, nF, 2, s.dinv, dmin);
        if (warp == 0) {
            T s12 = T(0), s22 = T(0);
            for (int k = lane; k < nF; k += 32) {
                const T y1 = Lb[k * LD + nF], y2 = Lb[k * LD + nF + 1], di = s.dinv[k];
                s12 = fma(y1 * y2, di, s12);
                s22 = fma(y2 * y2, di, s22);
            }
            s12 = warp_sum(s12); s22 = warp_sum(s22);
            const T nun = -s12 / s22;                   // 1' dw = 0
            T zt[NREG];
#pragma unroll
            for (int m = 0; m < NREG; ++m) {
                const int k = lane + 32 * m;
                zt[m] = k < nF ? -(Lb[k * LD + nF] + nun * Lb[k * LD + nF + 1]) : T(0);
            }
            backsub_warp<T, NREG>(Lb, LD, nF, s.dinv, zt);
            // step: decrement, the largest feasible step and the asset that blocks it at the cap
            T dec = T(0), dabs = T(0), wmax = T(0), a0 = T(1), au = Num<T>::inf();
            int iu = 0x7fffffff;
#pragma unroll
            for (int m = 0; m < NREG; ++m) {
                const int k = lane + 32 * m;
                if (k < nF) {
                    const int i = s.idx[k];
                    const T d = zt[m], wi = wv[i];
                    dw[k] = d;
                    dec = fma(-d, Qw[i] - bv[i] / wi + nun, dec);
                    dabs = max(dabs, absT(d) / wi); wmax = max(wmax, wi);     // relative step: the barrier scales every weight by itself
                    if (d < T(0)) a0 = min(a0, T(0.99) * (-wi / d));
                    if (d > T(0)) { const T r = (u - wi) / d; if (r < au) { au = r; iu = i; } }
                }
            }
            dec = warp_sum(dec); dabs = warp_max(dabs); wmax = warp_max(wmax); a0 = warp_min(a0);
            warp_argmin(au, iu);
            if (lane == 0) {
                s.red[32] = nun; s.red[33] = dec; s.red[34] = dabs; s.red[35] = wmax; s.red[36] = a0; s.red[37] = au;
                s.misc[1] = iu;
            }
        }
        __syncthreads();
        nu = s.red[32];
        const T dec = s.red[33], dabs = s.red[34], wmax = s.red[35];
        if (!(dabs == dabs) || !(dec == dec)) { nonfinite = 1; break; }
        // Newton converges quadratically and the rounding floor of a relative step is a few eps (the barrier's b / w^2 dominates
        // H_ii): 256 eps is above that floor and leaves the weights exact to ~1e-5 relative in float32, 1e-13 in float64
        if (dabs <= T(256) * eps || dec <= T(0)) {
            // stationary on the free set: release the capped asset with the most negative multiplier, if any
            if (warp == 0) {
                T mmin = Num<T>::inf();
                int im = 0x7fffffff;
                for (int i = lane; i < N; i += 32)
                    if (s.st[i] > 0) { const T mu = bv[i] / u - Qw[i] - nu; if (mu < mmin) { mmin = mu; im = i; } }
                warp_argmin(mmin, im);
                const T tolm = T(64) * eps * (absT(nu) + dmax * u + T(1));
                int rel = 0;
                if (lane == 0 && im < N && mmin < -tolm) { s.st[im] = 0; rel = 1; }
                if (lane == 0) s.misc[2] = rel;
            }
            __syncthreads();
            if (!s.misc[2]) { converged = 1; break; }
            continue;
        }
        // Q dw on the free columns, the scalars of the line search
        for (int i = tid; i < N; i += kThreads) {
            T acc = T(0);
            for (int m = 0; m < nF; ++m) acc = fma(sym_at(Qs, LD, i, s.idx[m]), dw[m], acc);
            Qd[i] = acc;
        }
        __syncthreads();
        if (warp == 0) {
            T wQw = T(0), wQd = T(0), dQd = T(0), bl0 = T(0);
            for (int i = lane; i < N; i += 32) { wQw = fma(wv[i], Qw[i], wQw); wQd = fma(wv[i], Qd[i], wQd); }
            for (int k = lane; k < nF; k += 32) {
                const int i = s.idx[k];
                dQd = fma(dw[k], Qd[i], dQd);
                bl0 = fma(bv[i], logT(wv[i]), bl0);
            }
            wQw = warp_sum(wQw); wQd = warp_sum(wQd); dQd = warp_sum(dQd); bl0 = warp_sum(bl0);
            const T f0 = T(0.5) * wQw - bl0;
            T a = min(s.red[36], s.red[37]);
            int block = s.red[37] <= s.red[36] ? s.misc[1] : -1;
            for (int ls = 0; ls < 50; ++ls) {
                T bl1 = T(0);
                for (int k = lane; k < nF; k += 32) {
                    const int i = s.idx[k];
                    const T wn = fma(a, dw[k], wv[i]);
                    bl1 = fma(bv[i], logT(wn > Num<T>::tiny() ? wn : Num<T>::tiny()), bl1);
                }
                bl1 = warp_sum(bl1);
                const T f1 = T(0.5) * (wQw + T(2) * a * wQd + a * a * dQd) - bl1;
                // below the resolution of f the test is meaningless: the Newton direction is a descent direction, take it
                if (f1 <= f0 - T(1e-4) * a * dec || a * dec <= T(16) * eps * (absT(f0) + T(1))) break;
                a *= T(0.5);
                block = -1;
            }
            for (int k = lane; k < nF; k += 32) { const int i = s.idx[k]; wv[i] = fma(a, dw[k], wv[i]); }
            __syncwarp();
            if (lane == 0 && block >= 0) { wv[block] = u; s.st[block] = 1; }
        }
        __syncthreads();
    }
    for (int i = tid; i < N; i += kThreads) {
        const T wi = s.st[i] > 0 ? u : clampT(wv[i], T(0), u);
        nonfinite |= !(wi == wi);
        w_out[i] = wi;
        st_out[i] = s.st[i] > 0 || wi >= u ? 1 : 0;
    }
    nonfinite = __syncthreads_or(nonfinite);
    if (tid == 0) {
        int st = lifted ? DDW_STATUS_REGULARIZED : 0;
        if (!converged || nonfinite) { st |= DDW_STATUS_INEXACT; atomicAdd(p.bad_count, 1); }
        p.status[b] = st | (iters << 16);        // bits 16..: Newton / active-set iterations used (informational)
    }
}

template <typename T, int kThreads, int NREG, bool kGlobal>
__global__ void __launch_bounds__(kThreads) risk_budgeting_bwd_kernel(RbParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.N, LD = p.LD, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;
    const long long b = blockIdx.x;
    MkSmem<T> s = carve<T, kGlobal>(smem_raw, N, LD, p.AS, p.CH, p.ws, b);
    const int NP = round_up(N + 2, 32);
    T *Qs = s.Qs, *Lb = s.Qs + 1;
    T *bv = s.r, *wv = s.wv, *gv = s.g;
    const T *Ab = p.A + b * (long long)N * N;
    gram_lower<T, kThreads, false>(Ab, nullptr, 0, 1, N, N, nullptr, nullptr, N, s.As, p.AS, p.CH, Qs, LD, T(0));
    T *dv = s.As, *pv = s.As + NP;          // d (full length) and A d; A w overwrites gv once d is known
    T dmax = T(0);
    for (int i = tid; i < N; i += kThreads) {
        const T bi = p.b[b * N + i];
        bv[i] = bi > Num<T>::tiny() ? bi : Num<T>::tiny();
        wv[i] = p.w_in[b * N + i];
        gv[i] = p.grad_w[b * N + i];
        s.st[i] = p.state_in[b * N + i];
        dv[i] = T(0);
        dmax = max(dmax, Qs[i * LD + i]);
    }
    dmax = warp_max(dmax);
    if (lane == 0) s.red[warp] = dmax;
    __syncthreads();
    dmax = T(0);
    for (int i = 0; i < kWarps; ++i) dmax = max(dmax, s.red[i]);
    const T dmin = T(4) * Num<T>::eps() * dmax + Num<T>::tiny();
    if (warp == 0) rb_free_list(s, N);
    __syncthreads();
    const int nF = s.misc[0];
    if (nF > 0) {
        for (int cc = warp; cc < nF; cc += kWarps) {
            const int jc = s.idx[cc];
            for (int a = cc + lane; a < nF; a += 32) {
                T v = Qs[s.idx[a] * LD + jc];
                if (a == cc) { const T wi = wv[jc]; v += bv[jc] / (wi * wi); }
                Lb[cc * LD + a] = v;
            }
        }
        for (int a = tid; a < nF; a += kThreads) { Lb[a * LD + nF] = gv[s.idx[a]]; Lb[a * LD + nF + 1] = T(1); }
        __syncthreads();
        factor_ldl<T, kThreads>(Lb, LD, nF, 2, s.dinv, dmin);
        if (warp == 0) {
            T s12 = T(0), s22 = T(0);
            for (int k = lane; k < nF; k += 32) {
                const T y1 = Lb[k * LD + nF], y2 = Lb[k * LD + nF + 1], di = s.dinv[k];
                s12 = fma(y1 * y2, di, s12);
                s22 = fma(y2 * y2, di, s22);
            }
            s12 = warp_sum(s12); s22 = warp_sum(s22);
            const T eta = s12 / s22;
            T zt[NREG];
#pragma unroll
            for (int m = 0; m < NREG; ++m) {
                const int k = lane + 32 * m;
                zt[m] = k < nF ? Lb[k * LD + nF] - eta * Lb[k * LD + nF + 1] : T(0);
            }
            backsub_warp<T, NREG>(Lb, LD, nF, s.dinv, zt);
#pragma unroll
            for (int m = 0; m < NREG; ++m) {
                const int k = lane + 32 * m;
                if (k < nF) dv[s.idx[k]] = zt[m];
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < N; i += kThreads) p.d_b[b * N + i] = s.st[i] == 0 ? dv[i] / wv[i] : T(0);
    // p = A d, q = A w: one warp per row of covmat_sqrt (coalesced)
    T *qv = gv;
    __syncthreads();
    for (int j = warp; j < N; j += kWarps) {
        T sp = T(0), sq = T(0);
        for (int k = lane; k < N; k += 32) {
            const T a = Ab[(long long)j * N + k];
            sp = fma(a, dv[k], sp); sq = fma(a, wv[k], sq);
        }
        sp = warp_sum(sp); sq = warp_sum(sq);
        if (lane == 0) { pv[j] = sp; qv[j] = sq; }
    }
    __syncthreads();
    T *dA = p.d_A + b * (long long)N * N;
    {
        int j = tid / N, k = tid - j * N;
        const int dj = kThreads / N, dk = kThreads % N;
        for (int e = tid; e < N * N; e += kThreads) {
            dA[e] = -fma(pv[j], wv[k], qv[j] * dv[k]);
            k += dk; j += dj;
            if (k >= N) { k -= N; ++j; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
struct RbPlan { int threads, nreg, LD, AS, CH; bool global_q; size_t smem; };

static RbPlan rb_plan(int N, int tsize) {
    RbPlan pl;
    pl.threads = N <= 128 ? 128 : (N <= 256 ? 256 : 512);
    pl.nreg = N <= 128 ? 4 : (N <= 256 ? 8 : (N <= 512 ? 16 : 32));
    pl.LD = (N + 3) | 1;
    pl.AS = round_up(N, 4);
    pl.CH = max(1, min(N, pl.threads * kPre / N));
    const int NP = round_up(N + 2, 32);
    while (2 * pl.CH * pl.AS < 2 * NP) ++pl.CH;       // the staging area doubles as two N-vectors after the Gram
    pl.global_q = false;
    pl.smem = mk_smem_bytes(N, pl.LD, pl.AS, pl.CH, tsize, false);
    if (pl.smem > (size_t)device_smem_limit()) {
        pl.global_q = true;
        pl.smem = mk_smem_bytes(N, pl.LD, pl.AS, pl.CH, tsize, true);
    }
    return pl;
}

size_t risk_budgeting_workspace_bytes(long long B, int N, int tsize) {
    const RbPlan pl = rb_plan(N, tsize);
    return 256 + (pl.global_q ? (size_t)B * N * pl.LD * tsize : 0);
}

template <typename T, int kThreads, int NREG, bool kGlobal>
static int rb_launch(const RbParams<T> &p, const RbPlan &pl, bool backward, cudaStream_t st) {
    auto kf = risk_budgeting_fwd_kernel<T, kThreads, NREG, kGlobal>;
    auto kb = risk_budgeting_bwd_kernel<T, kThreads, NREG, kGlobal>;
    static DeviceOnce configured;
    const int dev = current_device();
    if (configured.needs(dev)) {
        if (cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, device_smem_limit()) != cudaSuccess ||
            cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, device_smem_limit()) != cudaSuccess)
            return check_launch("cudaFuncSetAttribute(risk_budgeting)");
        configured.mark(dev);
    }
    if (backward) kb<<<(unsigned)p.B, kThreads, pl.smem, st>>>(p);
    else kf<<<(unsigned)p.B, kThreads, pl.smem, st>>>(p);
    return check_launch(backward ? "risk_budgeting_bwd_kernel" : "risk_budgeting_fwd_kernel");
}

template <typename T>
int risk_budgeting_t(bool backward, const void *A, const void *bud, const void *grad_w, const void *w_in, const int8_t *state_in,
                     double u, long long B, int N, void *w, int8_t *state, int32_t *status, void *d_A, void *d_b, void *ws,
                     size_t ws_bytes, cudaStream_t st) {
    if (!ws || ws_bytes < risk_budgeting_workspace_bytes(B, N, sizeof(T))) { set_error("ddw_risk_budgeting: workspace too small"); return DDW_ERR_WORKSPACE; }
    const RbPlan pl = rb_plan(N, sizeof(T));
    if (pl.smem > (size_t)device_smem_limit()) { set_error("ddw_risk_budgeting: n_assets=%d needs %zu B of shared memory", N, pl.smem); return DDW_ERR_UNSUPPORTED; }
    RbParams<T> p;
    memset(&p, 0, sizeof(p));
    p.A = (const T *)A; p.b = (const T *)bud; p.grad_w = (const T *)grad_w; p.w_in = (const T *)w_in; p.state_in = state_in;
    p.u = (T)u; p.B = B; p.N = N; p.LD = pl.LD; p.AS = pl.AS; p.CH = pl.CH;
    p.w = (T *)w; p.d_A = (T *)d_A; p.d_b = (T *)d_b; p.state = state; p.status = status;
    p.bad_count = reinterpret_cast<int *>(ws);       // first int of the workspace: portfolios left unsolved
    p.ws = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(ws) + 256);
    if (!backward) cudaMemsetAsync(ws, 0, 16, st);
    if (!pl.global_q) {
        if (pl.threads == 128) return rb_launch<T, 128, 4, false>(p, pl, backward, st);
        if (pl.threads == 256) return rb_launch<T, 256, 8, false>(p, pl, backward, st);
    }
    if (pl.threads == 128) return rb_launch<T, 128, 4, true>(p, pl, backward, st);
    if (pl.threads == 256) return rb_launch<T, 256, 8, true>(p, pl, backward, st);
    if (pl.nreg == 16) return rb_launch<T, 512, 16, true>(p, pl, backward, st);
    return rb_launch<T, 512, 32, true>(p, pl, backward, st);
}

template int risk_budgeting_t<float>(bool, const void *, const void *, const void *, const void *, const int8_t *, double, long long,
                                     int, void *, int8_t *, int32_t *, void *, void *, void *, size_t, cudaStream_t);
template int risk_budgeting_t<double>(bool, const void *, const void *, const void *, const void *, const int8_t *, double, long long,
                                      int, void *, int8_t *, int32_t *, void *, void *, void *, size_t, cudaStream_t);

}  // namespace ddw
