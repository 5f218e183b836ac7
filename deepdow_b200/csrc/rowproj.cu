// rowproj.cu -- row-wise allocators: capped-simplex projection (SparsemaxAllocator, reference
// deepdow/layers/allocate.py:554-560,580-595) and capped softmax (variational SoftmaxAllocator,
// allocate.py:470-475,495-514), forward and backward.
//
// One warp per row; the row lives in registers (NREG values per lane), every reduction is a
// warp shuffle, so a row costs exactly one coalesced read of x and one coalesced write of w:
// (2N+1) elements of HBM traffic forward, 3N backward.
#include "common.cuh"

namespace ddw {

constexpr int kRowWarps = 4;

template <typename T> struct RowParams {
    const T *x, *temperature, *grad_w, *w;
    T temperature_scalar, u;
    long long B;
    int N;
    T *out, *d_x, *d_t;
};

// ---- capped simplex: w = clip(v - tau, 0, u), sum w = 1 ----------------------------------------
template <typename T, int NREG>
__global__ void __launch_bounds__(kRowWarps * 32) capped_simplex_fwd_kernel(RowParams<T> p) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= p.B) return;
    const int N = p.N;
    const T u = p.u;
    const T t = p.temperature ? p.temperature[row] : p.temperature_scalar;
    const T *x = p.x + row * N;
    T v[NREG];
    T vmax = -Num<T>::inf(), vmin = Num<T>::inf();
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        const int i = lane + 32 * m;
        if (i < N) { v[m] = x[i] / t; vmax = v[m] > vmax ? v[m] : vmax; vmin = v[m] < vmin ? v[m] : vmin; }
        else v[m] = -Num<T>::inf();   // never in the support
    }
    vmax = warp_max(vmax); vmin = warp_min(vmin);
    // Root of the non-increasing piecewise-linear phi(tau) = sum clip(v - tau, 0, u) - 1 by a
    // bracketed Newton iteration: the Newton step is the exact solve for the current (free, capped)
    // sets, so the iteration ends at a fixed point; a step that leaves the bracket or fails to
    // halve the step before last is replaced by a bisection (the rtsafe rule).
    T lo = vmin - u, hi = vmax;   // phi(lo) = N u - 1 >= 0 > phi(hi) = -1
    T tau = vmax - (u < T(1) ? u : T(1));
    if (tau < lo) tau = lo;
    T dxold = hi - lo, dx = dxold;
#pragma unroll 1
    for (int it = 0; it < 200; ++it) {
        T sumF = T(0);
        int cntF = 0, cntU = 0;
#pragma unroll
        for (int m = 0; m < NREG; ++m) {
            const T dlt = v[m] - tau;
            if (dlt >= u) ++cntU;
            else if (dlt > T(0)) { ++cntF; sumF += v[m]; }
        }
        sumF = warp_sum(sumF);
        cntF = warp_sum(cntF); cntU = warp_sum(cntU);
        const T s = sumF - (T)cntF * tau + u * (T)cntU;
        if (s > T(1)) lo = tau; else hi = tau;
        bool ok = false;
        T nt = tau;
        if (cntF > 0) {
            nt = (sumF + u * (T)cntU - T(1)) / (T)cntF;
            if (nt == tau) break;
            ok = nt >= lo && nt <= hi && absT(T(2) * (nt - tau)) <= absT(dxold);
        } else if (absT(s - T(1)) <= T(4) * Num<T>::eps() * (T)N) break;
        dxold = dx;
        if (ok) dx = nt - tau;
        else { dx = T(0.5) * (hi - lo); nt = lo + dx; }
        if (nt == tau) break;
        tau = nt;
    }
    T *w = p.out + row * N;
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        const int i = lane + 32 * m;
        if (i < N) w[i] = clampT(v[m] - tau, T(0), u);
    }
}

// d_inp = g - mean_F(g) on F = {0 < w < u}, 0 elsewhere; d_x = d_inp / t; d_t = -sum(d_inp x) / t^2
template <typename T, int NREG>
__global__ void __launch_bounds__(kRowWarps * 32) capped_simplex_bwd_kernel(RowParams<T> p) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= p.B) return;
    const int N = p.N;
    const T u = p.u;
    const T t = p.temperature ? p.temperature[row] : p.temperature_scalar;
    const T *w = p.w + row * N, *g = p.grad_w + row * N, *x = p.x + row * N;
    T gv[NREG];
    bool fr[NREG];
    T sg = T(0);
    int cnt = 0;
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        const int i = lane + 32 * m;
        fr[m] = false; gv[m] = T(0);
        if (i < N) {
            const T wi = w[i];
            fr[m] = wi > T(0) && wi < u;
            if (fr[m]) { gv[m] = g[i]; sg += gv[m]; ++cnt; }
        }
    }
    sg = warp_sum(sg); cnt = warp_sum(cnt);
    const T mean = cnt > 0 ? sg / (T)cnt : T(0);
    T dt = T(0);
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        const int i = lane + 32 * m;
        if (i < N) {
            const T di = fr[m] ? gv[m] - mean : T(0);
            p.d_x[row * N + i] = di / t;
            if (p.d_t) dt = fma(di, x[i], dt);
        }
    }
    if (p.d_t) {
        dt = warp_sum(dt);
        if (lane == 0) p.d_t[row] = -dt / (t * t);
    }
}

// ---- capped softmax: w_i = min(u, exp(v_i - nu)), sum w = 1 ------------------------------------
template <typename T, int NREG>
__global__ void __launch_bounds__(kRowWarps * 32) capped_softmax_fwd_kernel(RowParams<T> p) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= p.B) return;
    const int N = p.N;
    const T u = p.u;
    const T t = p.temperature ? p.temperature[row] : p.temperature_scalar;
    const T *x = p.x + row * N;
    T e[NREG];
    T vmax = -Num<T>::inf();
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        const int i = lane + 32 * m;
        e[m] = i < N ? x[i] / t : -Num<T>::inf();
        vmax = e[m] > vmax ? e[m] : vmax;
    }
    vmax = warp_max(vmax);
#pragma unroll
    for (int m = 0; m < NREG; ++m) e[m] = (lane + 32 * m) < N ? expT(e[m] - vmax) : T(0);
    unsigned long long capped = 0ull;   // bit m: slot m of this lane sits at the cap
    int ncap = 0;
    T scale = T(0);
    for (int it = 0; it <= N; ++it) {
        T sF = T(0);
#pragma unroll
        for (int m = 0; m < NREG; ++m) sF += ((capped >> m) & 1ull) ? T(0) : e[m];
        sF = warp_sum(sF);
        const T mass = T(1) - u * (T)ncap;
        if (!(sF > T(0)) || !(mass > T(0))) { scale = T(0); break; }
        scale = mass / sF;
        int nnew = 0;
#pragma unroll
        for (int m = 0; m < NREG; ++m) {
            if (!((capped >> m) & 1ull) && (lane + 32 * m) < N && e[m] * scale > u) { capped |= 1ull << m; ++nnew; }
        }
        nnew = warp_sum(nnew);
        if (nnew == 0) break;
        ncap += nnew;
    }
    T *w = p.out + row * N;
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        const int i = lane + 32 * m;
        if (i < N) w[i] = ((capped >> m) & 1ull) ? u : e[m] * scale;
    }
}

// d_inp = w_F * (g - sum_F(w g) / sum_F(w)) on F = {w < u}; 0 on the capped assets
template <typename T, int NREG>
__global__ void __launch_bounds__(kRowWarps * 32) capped_softmax_bwd_kernel(RowParams<T> p) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= p.B) return;
    const int N = p.N;
    const T u = p.u;
    const T t = p.temperature ? p.temperature[row] : p.temperature_scalar;
    const T *w = p.w + row * N, *g = p.grad_w + row * N, *x = p.x + row * N;
    T wv[NREG], gv[NREG];
    T sw = T(0), swg = T(0);
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        const int i = lane + 32 * m;
        wv[m] = T(0); gv[m] = T(0);
        if (i < N) {
            const T wi = w[i];
            if (wi < u) { wv[m] = wi; gv[m] = g[i]; sw += wi; swg = fma(wi, gv[m], swg); }
        }
    }
    sw = warp_sum(sw); swg = warp_sum(swg);
    const T mean = sw > T(0) ? swg / sw : T(0);
    T dt = T(0);
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        const int i = lane + 32 * m;
        if (i < N) {
            const T di = wv[m] * (gv[m] - mean);
            p.d_x[row * N + i] = di / t;
            if (p.d_t) dt = fma(di, x[i], dt);
        }
    }
    if (p.d_t) {
        dt = warp_sum(dt);
        if (lane == 0) p.d_t[row] = -dt / (t * t);
    }
}

template <typename T, template <typename, int> class K> struct Dummy {};

#define DDW_ROW_DISPATCH(KERNEL)                                                              \
    do {                                                                                      \
        const unsigned grid = (unsigned)((p.B + kRowWarps - 1) / kRowWarps);                  \
        if (p.N <= 32) KERNEL<T, 1><<<grid, kRowWarps * 32, 0, st>>>(p);                      \
        else if (p.N <= 64) KERNEL<T, 2><<<grid, kRowWarps * 32, 0, st>>>(p);                 \
        else if (p.N <= 128) KERNEL<T, 4><<<grid, kRowWarps * 32, 0, st>>>(p);                \
        else if (p.N <= 256) KERNEL<T, 8><<<grid, kRowWarps * 32, 0, st>>>(p);                \
        else if (p.N <= 512) KERNEL<T, 16><<<grid, kRowWarps * 32, 0, st>>>(p);               \
        else if (p.N <= 1024) KERNEL<T, 32><<<grid, kRowWarps * 32, 0, st>>>(p);              \
        else if (p.N <= 2048) KERNEL<T, 64><<<grid, kRowWarps * 32, 0, st>>>(p);              \
        else { set_error(#KERNEL ": n_assets=%d > 2048 is not supported", p.N); return DDW_ERR_UNSUPPORTED; } \
        return check_launch(#KERNEL);                                                         \
    } while (0)

template <typename T> int capped_simplex_fwd_t(const RowParams<T> &p, cudaStream_t st) { DDW_ROW_DISPATCH(capped_simplex_fwd_kernel); }
template <typename T> int capped_simplex_bwd_t(const RowParams<T> &p, cudaStream_t st) { DDW_ROW_DISPATCH(capped_simplex_bwd_kernel); }
template <typename T> int capped_softmax_fwd_t(const RowParams<T> &p, cudaStream_t st) { DDW_ROW_DISPATCH(capped_softmax_fwd_kernel); }
template <typename T> int capped_softmax_bwd_t(const RowParams<T> &p, cudaStream_t st) { DDW_ROW_DISPATCH(capped_softmax_bwd_kernel); }

template <typename T>
int row_op_t(int which, const void *x, const void *temperature, double tscalar, double u, long long B, int N,
             const void *grad_w, const void *w, void *out, void *d_x, void *d_t, cudaStream_t st) {
    RowParams<T> p;
    p.x = (const T *)x; p.temperature = (const T *)temperature; p.grad_w = (const T *)grad_w; p.w = (const T *)w;
    p.temperature_scalar = (T)tscalar; p.u = (T)u; p.B = B; p.N = N;
    p.out = (T *)out; p.d_x = (T *)d_x; p.d_t = (T *)d_t;
    switch (which) {
        case 0: return capped_simplex_fwd_t<T>(p, st);
        case 1: return capped_simplex_bwd_t<T>(p, st);
        case 2: return capped_softmax_fwd_t<T>(p, st);
        default: return capped_softmax_bwd_t<T>(p, st);
    }
}

template int row_op_t<float>(int, const void *, const void *, double, double, long long, int, const void *,
                             const void *, void *, void *, void *, cudaStream_t);
template int row_op_t<double>(int, const void *, const void *, double, double, long long, int, const void *,
                              const void *, void *, void *, void *, cudaStream_t);

}  // namespace ddw
