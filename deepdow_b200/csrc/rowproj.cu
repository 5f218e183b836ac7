// rowproj.cu -- row-wise allocators: capped-simplex projection (SparsemaxAllocator, reference
// deepdow/layers/allocate.py:554-560,580-595) and capped softmax (variational SoftmaxAllocator,
// allocate.py:470-475,495-514), forward and backward.
//
// A row is owned by a GROUP of LPR lanes (4, 16 or 32) holding NREG = 16+ values each, so a warp works on
// 32 / LPR rows at once: a row of 100 is 4 lanes x 32 values, and a reduction over the row is 2 shuffle steps shared by
// eight rows instead of 5 for one (the kernels were bound by the SM's shuffle rate with one warp per row).  Integer counts
// travel packed in one word.  A row costs one coalesced read of x and one coalesced write of w: (2N+1) elements of HBM
// traffic forward, 3N backward.  Rows whose length is a multiple of 4 move as 16-byte vectors (row_load / row_store).
#include "common.cuh"
#include "linalg.cuh"   // st4

namespace ddw {

constexpr int kRowWarps = 4;

template <typename T> struct RowParams {
    const T *x, *temperature, *grad_w, *w;
    T temperature_scalar, u;
    long long B;
    int N;
    T *out, *d_x, *d_t;
};

// Element m of lane `sub` of a group.  Scalar layout: i = sub + LPR m (any n_assets).  Vector layout (n_assets % 4 == 0,
// 16-byte aligned rows): the lane owns 4 consecutive elements per 4 LPR-element block, i = 4 (sub + LPR (m / 4)) + m % 4,
// and moves them with one 16-byte access.
template <int LPR, bool kVec> __device__ __forceinline__ int row_elem(int sub, int m) {
    return kVec ? 4 * (sub + LPR * (m >> 2)) + (m & 3) : sub + LPR * m;
}
template <typename T, int NREG, int LPR, bool kVec>
__device__ __forceinline__ void row_load(const T *__restrict__ src, int N, int sub, T (&v)[NREG], T fill) {
    if (kVec) {
#pragma unroll
        for (int mv = 0; mv < NREG / 4; ++mv) {
            const int i = 4 * (sub + LPR * mv);
            if (i < N) {
                const Vec4<T> q = ld4(src + i);
#pragma unroll
                for (int c = 0; c < 4; ++c) v[4 * mv + c] = q.v[c];
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) v[4 * mv + c] = fill;
            }
        }
    } else {
#pragma unroll
        for (int m = 0; m < NREG; ++m) { const int i = sub + LPR * m; v[m] = i < N ? src[i] : fill; }
    }
}
template <typename T, int NREG, int LPR, bool kVec>
__device__ __forceinline__ void row_store(T *__restrict__ dst, int N, int sub, const T (&v)[NREG]) {
    if (kVec) {
#pragma unroll
        for (int mv = 0; mv < NREG / 4; ++mv) {
            const int i = 4 * (sub + LPR * mv);
            if (i < N) {
                Vec4<T> q;
#pragma unroll
                for (int c = 0; c < 4; ++c) q.v[c] = v[4 * mv + c];
                st4(dst + i, q);
            }
        }
    } else {
#pragma unroll
        for (int m = 0; m < NREG; ++m) { const int i = sub + LPR * m; if (i < N) dst[i] = v[m]; }
    }
}
// reductions over the LPR lanes of a group (xor offsets below LPR stay inside the aligned group); every lane of the
// warp executes them, so groups whose row is finished or out of range keep stepping with the others
template <int LPR, typename V> __device__ __forceinline__ V group_sum(V v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}
template <int LPR, typename V> __device__ __forceinline__ V group_max(V v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) { V t = __shfl_xor_sync(kFullMask, v, o); v = t > v ? t : v; }
    return v;
}
template <int LPR, typename V> __device__ __forceinline__ V group_min(V v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) { V t = __shfl_xor_sync(kFullMask, v, o); v = t < v ? t : v; }
    return v;
}

// v / t for the whole row from ONE division: q = v * (1/t) plus one residual correction, which is the correctly rounded
// quotient except at the ends of the exponent range (what the division's own fast path computes, without its checks)
__device__ __forceinline__ float div_by(float v, float t, float rt) { const float q = v * rt; return fmaf(fmaf(-q, t, v), rt, q); }
__device__ __forceinline__ double div_by(double v, double t, double) { return v / t; }

__device__ __forceinline__ float fminT(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double fminT(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float fmaxT(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double fmaxT(double a, double b) { return fmax(a, b); }
// 0 < c < u for a c already clipped to [0, u].  float: one unsigned compare on the bit patterns (c = 0 wraps to the top
// of the range), which keeps the unrolled loop free of long-lived predicates.
// quotient for a search step (any point of the bracket would do): the hardware reciprocal is enough
__device__ __forceinline__ float step_div(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double step_div(double a, double b) { return a / b; }
__device__ __forceinline__ unsigned inside_key(float u) { return (unsigned)__float_as_int(u) - 1u; }
__device__ __forceinline__ bool is_inside(float c, unsigned key) { return (unsigned)__float_as_int(c) - 1u < key; }
__device__ __forceinline__ double inside_key(double u) { return u; }
__device__ __forceinline__ bool is_inside(double c, double u) { return c > 0.0 && c < u; }

// The row of this lane's group; `live` is false for the groups past the end of the batch, which shadow the last row
// (same loads, no stores) so that the warp-wide shuffles stay convergent.  Returns false when the whole warp is past the end.
template <int LPR> __device__ __forceinline__ bool row_of_group(long long B, long long &row, int &sub, bool &live) {
    const int lane = threadIdx.x & 31;
    sub = lane & (LPR - 1);
    const long long first = ((long long)blockIdx.x * kRowWarps + (threadIdx.x >> 5)) * (32 / LPR);
    if (first >= B) return false;
    row = first + lane / LPR;
    live = row < B;
    if (!live) row = B - 1;
    return true;
}

// ---- capped simplex: w = clip(v - tau, 0, u), sum w = 1 ----------------------------------------
template <typename T, int NREG, int LPR, bool kVec>
__global__ void __launch_bounds__(kRowWarps * 32) capped_simplex_fwd_kernel(RowParams<T> p) {
    long long row; int sub; bool live;
    if (!row_of_group<LPR>(p.B, row, sub, live)) return;
    const int N = p.N;
    const T u = p.u;
    const T t = p.temperature ? p.temperature[row] : p.temperature_scalar;
    T v[NREG];
    row_load<T, NREG, LPR, kVec>(p.x + row * N, N, sub, v, T(0));
    T vmax = -Num<T>::inf(), vmin = Num<T>::inf();
    const bool unit_t = t == T(1);
    const T rt = T(1) / t;
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        const bool in = row_elem<LPR, kVec>(sub, m) < N;
        const T q = unit_t ? v[m] : div_by(v[m], t, rt);
        vmax = in && q > vmax ? q : vmax; vmin = in && q < vmin ? q : vmin;
        v[m] = in ? q : -Num<T>::inf();   // padding is never in the support
    }
    vmax = group_max<LPR>(vmax); vmin = group_min<LPR>(vmin);
    // Root of the non-increasing piecewise-linear phi(tau) = sum clip(v - tau, 0, u) - 1 inside a bracket
    // [lo, hi], phi(lo) >= 0 > phi(hi).  The Newton step is the exact solve for the current (free, capped) sets, so the
    // iteration ends at a fixed point; a Newton step that leaves the bracket or fails to halve the step before last (the
    // rtsafe rule) -- or does not exist because no asset is free at tau, the usual case when the cap binds and phi is a
    // staircase of flats and short ramps -- is replaced by the secant of the bracket ends (Illinois rule: an end kept
    // twice has its value halved), and by a bisection every other time once six steps have not been enough.
    T lo = vmin - u, hi = vmax;
    T flo = u * (T)N - T(1), fhi = T(-1);
    // start: without a binding cap the top asset alone gives phi(vmax - 1) >= 0 and Newton climbs monotonically from
    // there; with a cap, the secant of the initial bracket
    T tau = u >= T(1) ? vmax - T(1) : lo + (hi - lo) * (flo / (flo - fhi));
    if (!(tau >= lo)) tau = lo;
    if (tau > hi) tau = hi;
    T dxold = hi - lo, dx = dxold;
    int side = 0, nfall = 0;
    bool done = false;
    const auto ukey = inside_key(u);
    T c[NREG];   // clip(v - tau, 0, u) of the last evaluation: the weights once the row is done
#pragma unroll 1
    for (int it = 0; it < 200; ++it) {
        T s0 = T(0), s1 = T(0);
        int cntF = 0;   // assets strictly between the bounds = minus the slope of phi at tau
#pragma unroll
        for (int m = 0; m < NREG; ++m) {
            c[m] = fminT(fmaxT(v[m] - tau, T(0)), u);
            if (m & 1) s1 += c[m]; else s0 += c[m];
            cntF += is_inside(c[m], ukey) ? 1 : 0;
        }
        const T ssum = group_sum<LPR>(s0 + s1);
        cntF = group_sum<LPR>(cntF);
        if (!done) {
            const T f = ssum - T(1);
            if (f > T(0)) { lo = tau; flo = f; if (side == 1) fhi *= T(0.5); side = 1; }
            else { hi = tau; fhi = f; if (side == 2) flo *= T(0.5); side = 2; }
            bool ok = false;
            T nt = tau;
            if (cntF > 0) {
                nt = tau + step_div(f, (T)cntF);
                if (nt == tau) done = true;
                ok = nt > lo && nt < hi && absT(T(2) * (nt - tau)) <= absT(dxold);
            } else if (absT(f) <= T(4) * Num<T>::eps() * (T)N) done = true;
            if (!done) {
                dxold = dx;
                if (!ok) {
                    const T den = flo - fhi;
                    nt = lo + (hi - lo) * step_div(flo, den);
                    const bool bisect = (it >= 6 && (nfall & 1)) || !(den > T(0)) || !(nt > lo && nt < hi);
                    if (bisect) nt = lo + T(0.5) * (hi - lo);
                    ++nfall;
                    // no number left strictly inside the bracket: tau is within one rounding of the root
                    if (!(nt > lo && nt < hi)) { done = true; nt = tau; }
                }
                dx = nt - tau;
                tau = nt;
            }
        }
        if (__all_sync(kFullMask, done)) break;
    }
    // Polish: the search steps above take phi from a sum of clipped values (rounding of order eps) and a hardware
    // reciprocal; the last step is the exact solve for the sets found, tau = (sum_F v + u |U| - 1) / |F| with a true
    // division, which lands exactly on a breakpoint when the solution sits on one (supports match the oracle's).
    {
        T sumF = T(0);
        int cnt = 0;   // free count in the low half, capped count in the high half (n_assets <= 2048)
#pragma unroll
        for (int m = 0; m < NREG; ++m) {
            const T cm = fminT(fmaxT(v[m] - tau, T(0)), u);
            const bool fre = is_inside(cm, ukey);
            sumF += fre ? v[m] : T(0);
            cnt += fre ? 1 : (cm > T(0) ? (1 << 16) : 0);
        }
        sumF = group_sum<LPR>(sumF);
        cnt = group_sum<LPR>(cnt);
        const int cntF = cnt & 0xffff, cntU = cnt >> 16;
        if (cntF > 0) {
            const T tp = (sumF + (u * (T)cntU - T(1))) / (T)cntF;
            if (absT(tp - tau) <= T(1e-3) * (absT(tau) + T(1))) tau = tp;
        }
#pragma unroll
        for (int m = 0; m < NREG; ++m) c[m] = fminT(fmaxT(v[m] - tau, T(0)), u);
    }
    if (live) row_store<T, NREG, LPR, kVec>(p.out + row * N, N, sub, c);
}

// d_inp = g - mean_F(g) on F = {0 < w < u}, 0 elsewhere; d_x = d_inp / t; d_t = -sum(d_inp x) / t^2
template <typename T, int NREG, int LPR, bool kVec>
__global__ void __launch_bounds__(kRowWarps * 32) capped_simplex_bwd_kernel(RowParams<T> p) {
    long long row; int sub; bool live;
    if (!row_of_group<LPR>(p.B, row, sub, live)) return;
    const int N = p.N;
    const T u = p.u;
    const T t = p.temperature ? p.temperature[row] : p.temperature_scalar;
    T wv[NREG], gv[NREG];
    row_load<T, NREG, LPR, kVec>(p.w + row * N, N, sub, wv, T(0));
    row_load<T, NREG, LPR, kVec>(p.grad_w + row * N, N, sub, gv, T(0));
    bool fr[NREG];
    T sg = T(0);
    int cnt = 0;
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        fr[m] = row_elem<LPR, kVec>(sub, m) < N && wv[m] > T(0) && wv[m] < u;
        gv[m] = fr[m] ? gv[m] : T(0);
        sg += gv[m]; cnt += fr[m] ? 1 : 0;
    }
    sg = group_sum<LPR>(sg); cnt = group_sum<LPR>(cnt);
    const T mean = cnt > 0 ? sg / (T)cnt : T(0);
    const bool unit_t = t == T(1);
    const T rt = T(1) / t;
    T dt = T(0);
    if (p.d_t) row_load<T, NREG, LPR, kVec>(p.x + row * N, N, sub, wv, T(0));   // w is not needed any more
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        const T di = fr[m] ? gv[m] - mean : T(0);
        if (p.d_t) dt = fma(di, wv[m], dt);
        gv[m] = unit_t ? di : div_by(di, t, rt);
    }
    if (live) row_store<T, NREG, LPR, kVec>(p.d_x + row * N, N, sub, gv);
    if (p.d_t) {
        dt = group_sum<LPR>(dt);
        if (live && sub == 0) p.d_t[row] = -dt / (t * t);
    }
}

// ---- capped softmax: w_i = min(u, exp(v_i - nu)), sum w = 1 ------------------------------------
template <typename T, int NREG, int LPR, bool kVec>
__global__ void __launch_bounds__(kRowWarps * 32) capped_softmax_fwd_kernel(RowParams<T> p) {
    long long row; int sub; bool live;
    if (!row_of_group<LPR>(p.B, row, sub, live)) return;
    const int N = p.N;
    const T u = p.u;
    const T t = p.temperature ? p.temperature[row] : p.temperature_scalar;
    T v[NREG], e[NREG];
    row_load<T, NREG, LPR, kVec>(p.x + row * N, N, sub, v, T(0));
    T vmax = -Num<T>::inf();
    const bool unit_t = t == T(1);
    const T rt = T(1) / t;
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        const T q = unit_t ? v[m] : div_by(v[m], t, rt);
        v[m] = row_elem<LPR, kVec>(sub, m) < N ? q : -Num<T>::inf();
        vmax = v[m] > vmax ? v[m] : vmax;
    }
    vmax = group_max<LPR>(vmax);
#pragma unroll
    for (int m = 0; m < NREG; ++m) e[m] = row_elem<LPR, kVec>(sub, m) < N ? expT(v[m] - vmax) : T(0);
    // The capped set only grows: rescale the rest to the remaining mass, cap what now exceeds u, repeat.  This is Newton on
    // the concave piecewise-linear sum min(u, e s) = 1 in the scale s, from the left: monotone, ends after finitely many steps.
    // When the logits are spread wider than the exponent range and the cap pushes mass far down the row, the exponentials
    // of the assets still free underflow against the row maximum: they are then re-based on the largest free logit.
    const T tiny_sum = sizeof(T) == 4 ? T(1e-30) : T(1e-290);
    int ncap = 0;                      // a capped slot is marked by e = -1 (padding holds 0 and never exceeds the cap)
    T scale = T(0);
    bool done = false;
#pragma unroll 1
    for (int it = 0; it <= N; ++it) {
        T sF = T(0);
#pragma unroll
        for (int m = 0; m < NREG; ++m) sF += e[m] > T(0) ? e[m] : T(0);
        sF = group_sum<LPR>(sF);
        const bool rebase = !done && sF < tiny_sum && ncap < N;
        if (__any_sync(kFullMask, rebase)) {
            T vm = -Num<T>::inf();
#pragma unroll
            for (int m = 0; m < NREG; ++m) vm = (e[m] >= T(0) && v[m] > vm) ? v[m] : vm;
            vm = group_max<LPR>(vm);
            T s2 = T(0);
#pragma unroll
            for (int m = 0; m < NREG; ++m) {
                if (rebase && e[m] >= T(0)) e[m] = expT(v[m] - vm);      // padding: exp(-inf) = 0
                s2 += e[m] > T(0) ? e[m] : T(0);
            }
            sF = group_sum<LPR>(s2);
        }
        const T mass = T(1) - u * (T)ncap;
        const bool empty = !(sF > T(0)) || !(mass > T(0));
        if (!done) scale = empty ? T(0) : mass / sF;
        int nnew = 0;
#pragma unroll
        for (int m = 0; m < NREG; ++m) {
            const bool c = !done && e[m] * scale > u;
            e[m] = c ? T(-1) : e[m];
            nnew += c ? 1 : 0;
        }
        nnew = group_sum<LPR>(nnew);
        if (nnew == 0) done = true;
        ncap += nnew;
        if (__all_sync(kFullMask, done)) break;
    }
#pragma unroll
    for (int m = 0; m < NREG; ++m) e[m] = e[m] < T(0) ? u : e[m] * scale;
    if (live) row_store<T, NREG, LPR, kVec>(p.out + row * N, N, sub, e);
}

// d_inp = w_F * (g - sum_F(w g) / sum_F(w)) on F = {w < u}; 0 on the capped assets
template <typename T, int NREG, int LPR, bool kVec>
__global__ void __launch_bounds__(kRowWarps * 32) capped_softmax_bwd_kernel(RowParams<T> p) {
    long long row; int sub; bool live;
    if (!row_of_group<LPR>(p.B, row, sub, live)) return;
    const int N = p.N;
    const T u = p.u;
    const T t = p.temperature ? p.temperature[row] : p.temperature_scalar;
    T wv[NREG], gv[NREG];
    row_load<T, NREG, LPR, kVec>(p.w + row * N, N, sub, wv, u);          // fill = u: padding counts as capped
    row_load<T, NREG, LPR, kVec>(p.grad_w + row * N, N, sub, gv, T(0));
    T sw = T(0), swg = T(0);
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        const bool fre = row_elem<LPR, kVec>(sub, m) < N && wv[m] < u;
        wv[m] = fre ? wv[m] : T(0);
        sw += wv[m]; swg = fma(wv[m], gv[m], swg);
    }
    sw = group_sum<LPR>(sw); swg = group_sum<LPR>(swg);
    const T mean = sw > T(0) ? swg / sw : T(0);
    const bool unit_t = t == T(1);
    const T rt = T(1) / t;
    T dt = T(0);
#pragma unroll
    for (int m = 0; m < NREG; ++m) wv[m] = wv[m] * (gv[m] - mean);       // d_inp
    if (p.d_t) {
        row_load<T, NREG, LPR, kVec>(p.x + row * N, N, sub, gv, T(0));   // g is not needed any more
#pragma unroll
        for (int m = 0; m < NREG; ++m) dt = fma(wv[m], gv[m], dt);
    }
#pragma unroll
    for (int m = 0; m < NREG; ++m) wv[m] = unit_t ? wv[m] : div_by(wv[m], t, rt);
    if (live) row_store<T, NREG, LPR, kVec>(p.d_x + row * N, N, sub, wv);
    if (p.d_t) {
        dt = group_sum<LPR>(dt);
        if (live && sub == 0) p.d_t[row] = -dt / (t * t);
    }
}

static bool aligned16(const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

// 16 values per lane and as many lanes per row as that takes (4 ... 32); longer rows keep the full warp and grow NREG
#define DDW_ROW_LAUNCH(KERNEL, NREG, LPR, VEC)                                                            \
    do {                                                                                                  \
        const long long rows_per_block = (long long)kRowWarps * (32 / (LPR));                             \
        const unsigned grid = (unsigned)((p.B + rows_per_block - 1) / rows_per_block);                    \
        KERNEL<T, NREG, LPR, VEC><<<grid, kRowWarps * 32, 0, st>>>(p);                                    \
    } while (0)
#define DDW_ROW_SIZES(KERNEL, VEC)                                                                        \
    do {                                                                                                  \
        if (p.N <= 64) DDW_ROW_LAUNCH(KERNEL, 16, 4, VEC);                                                \
        else if (p.N <= 128) DDW_ROW_LAUNCH(KERNEL, 32, 4, VEC);                                          \
        else if (p.N <= 256) DDW_ROW_LAUNCH(KERNEL, 16, 16, VEC);                                         \
        else if (p.N <= 512) DDW_ROW_LAUNCH(KERNEL, 16, 32, VEC);                                         \
        else if (p.N <= 1024) DDW_ROW_LAUNCH(KERNEL, 32, 32, VEC);                                        \
        else if (p.N <= 2048) DDW_ROW_LAUNCH(KERNEL, 64, 32, VEC);                                        \
        else { set_error(#KERNEL ": n_assets=%d > 2048 is not supported", p.N); return DDW_ERR_UNSUPPORTED; } \
    } while (0)
#define DDW_ROW_DISPATCH(KERNEL)                                                                          \
    do {                                                                                                  \
        if (p.B <= 0) return 0;                                                                           \
        /* vector layout: rows are multiples of 16 bytes and every array starts on one */                  \
        const bool vec = (p.N % 4 == 0) && aligned16(p.x) && aligned16(p.grad_w) && aligned16(p.w) &&     \
                         aligned16(p.out) && aligned16(p.d_x);                                            \
        if (vec) DDW_ROW_SIZES(KERNEL, true); else DDW_ROW_SIZES(KERNEL, false);                          \
        return check_launch(#KERNEL);                                                                     \
    } while (0)

template <typename T> int capped_simplex_fwd_t(const RowParams<T> &p, cudaStream_t st) { DDW_ROW_DISPATCH(capped_simplex_fwd_kernel); }
template <typename T> int capped_simplex_bwd_t(const RowParams<T> &p, cudaStream_t st) { DDW_ROW_DISPATCH(capped_simplex_bwd_kernel); }
template <typename T> int capped_softmax_fwd_t(const RowParams<T> &p, cudaStream_t st) { DDW_ROW_DISPATCH(capped_softmax_fwd_kernel); }
template <typename T> int capped_softmax_bwd_t(const RowParams<T> &p, cudaStream_t st) { DDW_ROW_DISPATCH(capped_softmax_bwd_kernel); }


// ---- box-simplex LP: argmax x'w, sum w = 1, 0 <= w <= u (MaximumReturn, deepdow/benchmarks.py:114-125,160) -------------
// The optimum fills the cap greedily from the largest entry down: floor(1/u) assets at u, the next one takes the rest.
// One selection round per funded asset: group-wide (value, index) arg-max, ties to the smaller index.
template <typename T, int NREG, int LPR, bool kVec>
__global__ void __launch_bounds__(kRowWarps * 32) capped_argmax_kernel(RowParams<T> p, int kfull, T rem) {
    long long row; int sub; bool live;
    if (!row_of_group<LPR>(p.B, row, sub, live)) return;
    const int N = p.N;
    T v[NREG], w[NREG];
    row_load<T, NREG, LPR, kVec>(p.x + row * N, N, sub, v, T(0));
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        w[m] = T(0);
        if (row_elem<LPR, kVec>(sub, m) >= N) v[m] = -Num<T>::inf();
        else if (!(v[m] == v[m])) v[m] = -Num<T>::inf();      // NaN never wins a round
    }
    const int rounds = kfull + (rem > T(0) ? 1 : 0);
#pragma unroll 1
    for (int r = 0; r < rounds && r < N; ++r) {
        T best = -Num<T>::inf();
        int bi = 0x7fffffff;
#pragma unroll
        for (int m = 0; m < NREG; ++m) {
            const int i = row_elem<LPR, kVec>(sub, m);
            const bool open = i < N && w[m] == T(0);
            if (open && (v[m] > best || (v[m] == best && i < bi))) { best = v[m]; bi = i; }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
            const T ob = __shfl_xor_sync(kFullMask, best, o);
            const int oi = __shfl_xor_sync(kFullMask, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        const T give = r < kfull ? p.u : rem;
#pragma unroll
        for (int m = 0; m < NREG; ++m)
            if (row_elem<LPR, kVec>(sub, m) == bi) w[m] = give;
    }
    if (live) row_store<T, NREG, LPR, kVec>(p.out + row * N, N, sub, w);
}

template <typename T> int capped_argmax_t(const RowParams<T> &p, cudaStream_t st) {
    // funded assets at the cap, and what is left for the next one (computed in double on the host)
    const double u = (double)p.u;
    int kfull = (int)floor(1.0 / u + 1e-9);
    if (kfull > p.N) kfull = p.N;
    double remd = 1.0 - kfull * u;
    if (remd < 1e-12 || kfull >= p.N) remd = 0.0;
    const T rem = (T)remd;
#define DDW_ARGMAX_LAUNCH(NREG, LPR, VEC)                                                                 \
    do {                                                                                                  \
        const long long rows_per_block = (long long)kRowWarps * (32 / (LPR));                             \
        const unsigned grid = (unsigned)((p.B + rows_per_block - 1) / rows_per_block);                    \
        capped_argmax_kernel<T, NREG, LPR, VEC><<<grid, kRowWarps * 32, 0, st>>>(p, kfull, rem);          \
    } while (0)
#define DDW_ARGMAX_SIZES(VEC)                                                                             \
    do {                                                                                                  \
        if (p.N <= 64) DDW_ARGMAX_LAUNCH(16, 4, VEC);                                                     \
        else if (p.N <= 128) DDW_ARGMAX_LAUNCH(16, 8, VEC);                                               \
        else if (p.N <= 256) DDW_ARGMAX_LAUNCH(16, 16, VEC);                                              \
        else if (p.N <= 512) DDW_ARGMAX_LAUNCH(16, 32, VEC);                                              \
        else if (p.N <= 1024) DDW_ARGMAX_LAUNCH(32, 32, VEC);                                             \
        else if (p.N <= 2048) DDW_ARGMAX_LAUNCH(64, 32, VEC);                                             \
        else { set_error("capped_argmax_kernel: n_assets=%d > 2048 is not supported", p.N); return DDW_ERR_UNSUPPORTED; } \
    } while (0)
    if (p.B <= 0) return 0;
    const bool vec = (p.N % 4 == 0) && aligned16(p.x) && aligned16(p.out);
    if (vec) DDW_ARGMAX_SIZES(true); else DDW_ARGMAX_SIZES(false);
    return check_launch("capped_argmax_kernel");
#undef DDW_ARGMAX_SIZES
#undef DDW_ARGMAX_LAUNCH
}

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
        case 4: return capped_argmax_t<T>(p, st);
        default: return capped_softmax_bwd_t<T>(p, st);
    }
}

template int row_op_t<float>(int, const void *, const void *, double, double, long long, int, const void *,
                             const void *, void *, void *, void *, cudaStream_t);
template int row_op_t<double>(int, const void *, const void *, double, double, long long, int, const void *,
                              const void *, void *, void *, void *, cudaStream_t);

}  // namespace ddw
