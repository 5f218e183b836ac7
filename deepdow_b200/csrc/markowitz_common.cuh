// markowitz_common.cuh -- pieces shared by the dense and the sparse-support NumericalMarkowitz forward kernels:
// launch parameters, the small in-register factorisation, one primal-dual active-set step on a dense Q,
// the diagonal warm start and the result writers.  See markowitz_impl.cuh for the problem statement.
#pragma once
#include "common.cuh"
#include "linalg.cuh"
#include "cov_tile.cuh"

namespace ddw {

// Optional phase timing (build with -DDDW_PHASE_TIMING, tools/phase_timing.py): thread 0 of every CTA adds the
// cycles it spent between phase marks to a global table.  Compiled out of the product build.
#ifdef DDW_PHASE_TIMING
static __device__ unsigned long long ddw_phase_cycles[16];
#define DDW_PHASE_INIT long long ph_t0 = clock64();
#define DDW_PHASE(id) do { if (threadIdx.x == 0) { const long long ph_t1 = clock64(); atomicAdd(&ddw_phase_cycles[id], (unsigned long long)(ph_t1 - ph_t0)); ph_t0 = ph_t1; } } while (0)
#define DDW_PHASE_COUNT(id) do { if (threadIdx.x == 0) atomicAdd(&ddw_phase_cycles[id], 1ull); } while (0)
#define DDW_PHASE_ADD(id, v) do { if (threadIdx.x == 0) atomicAdd(&ddw_phase_cycles[id], (unsigned long long)(v)); } while (0)
#define DDW_SUBPHASE_INIT long long sub_t0 = clock64();
#define DDW_SUBPHASE(id) do { if (threadIdx.x == 0) { const long long sub_t1 = clock64(); atomicAdd(&ddw_phase_cycles[id], (unsigned long long)(sub_t1 - sub_t0)); sub_t0 = sub_t1; } } while (0)
#else
#define DDW_SUBPHASE_INIT
#define DDW_SUBPHASE(id) do { } while (0)
#define DDW_PHASE_INIT
#define DDW_PHASE(id) do { } while (0)
#define DDW_PHASE_COUNT(id) do { } while (0)
#define DDW_PHASE_ADD(id, v) do { } while (0)
#endif


constexpr int kMaxPdas = 24;   // active-set iterations before the fallback is asked for
constexpr int kMaxIpm = 40;

template <typename T> struct MkFwdParams {
    const T *rets, *C, *gamma, *alpha;
    T u;
    long long B;       // problems in this launch
    long long Bmod;    // n_samples of the whole batch = modulus of the gamma tiling
    long long b0;      // index of this launch's first problem in the whole batch
    int N, LD, AS, CH;
    T *w;
    int8_t *state;
    int32_t *status;
    T *ws;            // global Q/L storage when the matrix does not fit shared memory
    int *fail_count;  // [0] = number of problems that need the fallback
    int *fail_list;   // their indices
    int *work_count;  // dense kernel in list mode: number of entries of work_list (nullptr: one CTA per problem)
    int *work_list;
    int *bad_count;   // number of portfolios left INFEASIBLE or INEXACT (what the host layer raises on)
    // optional: the factor of the converged reduced KKT system, kept for the backward (ddw_markowitz_fwd_saved)
    int *saved_n;            // per problem: number of free assets whose factor was saved, -1 = none
    T *saved;                // Q in shared memory: packed columns + 1 / pivots, `saved_stride` elements per problem;
    long long saved_stride;  // Q in the global workspace: the workspace itself (the kernel works in place in it)
    // fused covariance -> solver (ddw_covariance_markowitz_fwd): covmat_sqrt is not read from HBM but formed in shared memory
    // from the (L, N) lookback tile as CovarianceMatrix(sqrt=False) would (BachelierNet, nn.py:170-171,189-191)
    const T *X, *coef;
    int L, strategy;
};

template <typename T> struct MkBwdParams {
    const T *grad_w, *w, *C, *gamma, *alpha;
    const int8_t *state;
    long long B, Bmod, b0;   // as in MkFwdParams
    int N, LDW, AS, CH;
    T *d_rets, *d_C, *d_alpha, *vecs;
    T *ws;  // global factor storage when it does not fit shared memory
    int *work_count, *work_list;   // sparse-support backward -> dense backward hand-over (nullptr: one CTA per problem)
    // optional: factors saved by the forward (ddw_markowitz_bwd_saved); problems without one go to work_list2
    const int *saved_n;
    const T *saved;
    long long saved_stride;
    int LDF;                       // row stride of the forward's Q / factor array (global variant of `saved`)
    int *work_count2, *work_list2;
};

template <typename T> __device__ __forceinline__ T sym_at(const T *Qs, int LD, int i, int j) {
    return i >= j ? Qs[i * LD + j] : Qs[j * LD + i];
}

// Shared-memory carve-up used by the forward kernels
template <typename T> struct MkSmem {
    T *Qs, *As, *r, *wv, *g, *dinv, *red;
    int *idx, *idxU, *misc;
    int8_t *st;
};

template <typename T, bool kGlobal>
__device__ __forceinline__ MkSmem<T> carve(unsigned char *raw, int N, int LD, int AS, int CH, T *ws, long long b) {
    MkSmem<T> s;
    const int NP = round_up(N + 2, 32);
    T *p = reinterpret_cast<T *>(raw);
    if (kGlobal) s.Qs = ws + b * (long long)N * LD;
    else { s.Qs = p; p += round_up(N * LD, 4); }
    s.As = p; p += 2 * CH * AS;
    s.r = p; p += NP;
    s.wv = p; p += NP;
    s.g = p; p += NP;
    s.dinv = p; p += NP;
    s.red = p; p += 64;
    int *q = reinterpret_cast<int *>(p);
    s.idx = q; q += NP;
    s.idxU = q; q += NP;
    s.misc = q; q += 8;
    s.st = reinterpret_cast<int8_t *>(q);
    return s;
}

__host__ __device__ static inline size_t mk_smem_bytes(int N, int LD, int AS, int CH, int tsize, bool global_q) {
    const int NP = round_up(N + 2, 32);
    size_t elems = (global_q ? 0 : (size_t)round_up(N * LD, 4)) + (size_t)2 * CH * AS + (size_t)4 * NP + 64;
    return elems * tsize + ((size_t)2 * NP + 8) * sizeof(int) + (size_t)NP;
}

// ------------------------------------------------------------------------------------------------
// Small reduced systems (the usual case: a handful of free assets): every element of the augmented
// lower triangle (n columns, n + 2 rows: Q_FF on top of the two right-hand sides) lives in a
// register of its owning thread for the whole factorisation.  Column k is published to shared
// memory once it is final; all later columns then apply  v(i,j) -= L(i,k) L(j,k) / d_k  from
// registers.  One barrier per column, no index arithmetic in the loop.
// ------------------------------------------------------------------------------------------------
constexpr int kSmallElemsMax = 8;   // elements per thread (variants with 2, 4 and 8 slots)

// number of register slots the augmented triangle of an n x n system needs: the matrix part is enumerated as a
// folded rectangle (row r of the rectangle = matrix row n-1-r followed by matrix row r, n+1 entries), the two
// right-hand-side rows follow.  The fold makes element -> (i,j) two multiplications instead of a square root.
__host__ __device__ __forceinline__ int small_system_elems(int n) { return ((n + 1) >> 1) * (n + 1) + 2 * n; }

template <typename T, int kThreads>
__device__ __forceinline__ bool small_system_fits(int n) { return small_system_elems(n) <= kThreads * kSmallElemsMax; }

template <typename T, int kThreads, int kSmallElems, typename Init>
__device__ __forceinline__ int factor_small_n(T *Lb, int ld, int nF, T *dinv, T dmin, Init init) {
    const int tid = threadIdx.x;
    const int np1 = nF + 1, EM = ((nF + 1) >> 1) * np1, E = EM + 2 * nF;
    const float inv_np1 = 1.f / (float)np1, inv_n = 1.f / (float)nF;
    T v[kSmallElems];
    int ii[kSmallElems], jj[kSmallElems];
#pragma unroll
    for (int m = 0; m < kSmallElems; ++m) {
        const int e = tid + m * kThreads;
        ii[m] = 0; jj[m] = -1; v[m] = T(0);
        if (e < E) {
            int i, j;
            bool valid = true;
            if (e < EM) {
                const int r = (int)(((float)e + 0.5f) * inv_np1), c = e - r * np1;   // exact: e < 2^11, margin 0.5/np1
                if (c < nF - r) { i = nF - 1 - r; j = c; }
                else { i = r; j = c - (nF - r); valid = (r != nF - 1 - r); }             // odd n: the middle row folds onto itself
            } else {
                const int e2 = e - EM, q = (int)(((float)e2 + 0.5f) * inv_n);
                i = nF + q; j = e2 - q * nF;
            }
            if (valid) { ii[m] = i; jj[m] = j; v[m] = init(i, j); }
        }
    }
    int lifted = 0;
    for (int k = 0; k < nF; ++k) {
        T *colk = Lb + k * ld;
#pragma unroll
        for (int m = 0; m < kSmallElems; ++m)
            if (jj[m] == k) colk[ii[m]] = v[m];
        __syncthreads();
        // the column entries first, the pivot's reciprocal second: the loads do not wait for the reciprocal chain
        T ci[kSmallElems], cj[kSmallElems];
#pragma unroll
        for (int m = 0; m < kSmallElems; ++m) {
            const bool act = jj[m] > k;
            ci[m] = act ? colk[ii[m]] : T(0);
            cj[m] = act ? colk[jj[m]] : T(0);
        }
        T d = colk[k];
        if (!(d > dmin)) { d = dmin; lifted = 1; }
        const T invd = recipFastT(d);
        if (tid == 0) dinv[k] = invd;
#pragma unroll
        for (int m = 0; m < kSmallElems; ++m) v[m] = fma(-ci[m], cj[m] * invd, v[m]);
    }
    __syncthreads();
    return lifted;
}

// init(i, j) returns element (i, j) of the augmented lower triangle: rows 0..nF-1 the matrix, row nF
// and nF+1 the two right-hand sides.
template <typename T, int kThreads, typename Init>
__device__ __forceinline__ int factor_small(T *Lb, int ld, int nF, T *dinv, T dmin, Init init) {
    const int E = small_system_elems(nF);
    if (E <= 2 * kThreads) return factor_small_n<T, kThreads, 2>(Lb, ld, nF, dinv, dmin, init);
    if (E <= 4 * kThreads) return factor_small_n<T, kThreads, 4>(Lb, ld, nF, dinv, dmin, init);
    return factor_small_n<T, kThreads, 8>(Lb, ld, nF, dinv, dmin, init);
}

// ------------------------------------------------------------------------------------------------
// One reduced-KKT solve + state update (a primal-dual active-set step).  Returns `changed`.
//   w_L = 0, w_U = u, [[2 Q_FF, 1],[1', 0]] [w_F; nu] = [r_F - 2 u Q_FU 1; 1 - u |U|]
// ------------------------------------------------------------------------------------------------
template <typename T, int kThreads, int NREG>
__device__ __forceinline__ int pdas_step(const MkSmem<T> &s, int N, int LD, T u, T dmin, T tolw, int &lifted) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;
    T *Qs = s.Qs, *Lb = s.Qs + 1;
    // (a) ordered index lists of the free and the capped assets
    if (warp == 0) {
        int nf = 0, nu = 0;
        for (int base = 0; base < N; base += 32) {
            const int i = base + lane;
            const int v = i < N ? (int)s.st[i] : 2;
            const unsigned mf = __ballot_sync(kFullMask, v == 0), mu = __ballot_sync(kFullMask, v == 1);
            const unsigned lt = (1u << lane) - 1u;
            if (v == 0) s.idx[nf + __popc(mf & lt)] = i;
            if (v == 1) s.idxU[nu + __popc(mu & lt)] = i;
            nf += __popc(mf); nu += __popc(mu);
        }
        if (lane == 0) { s.misc[0] = nf; s.misc[1] = nu; }
    }
    for (int i = tid; i < N; i += kThreads) s.wv[i] = s.st[i] > 0 ? u : T(0);
    __syncthreads();
    const int nF = s.misc[0], nU = s.misc[1];
    const T c = T(1) - u * (T)nU;
    int changed = 0;
    T nu_mult = T(0);
    if (nF > 0) {
        if (small_system_fits<T, kThreads>(nF)) {
            // (b,c) gather straight into registers and factor there
            auto init = [&](int i, int j) -> T {
                const int cj = s.idx[j];
                if (i < nF) return Qs[s.idx[i] * LD + cj];   // idx ascending: (idx[i], idx[j]) is in the lower triangle
                if (i > nF) return T(1);
                T acc = T(0);
                for (int q = 0; q < nU; ++q) acc += sym_at(Qs, LD, cj, s.idxU[q]);
                return s.r[cj] - T(2) * u * acc;
            };
            lifted |= factor_small<T, kThreads>(Lb, LD, nF, s.dinv, dmin, init);
        } else {
            // (b) gather Q_FF and the two right-hand sides
            for (int cc = warp; cc < nF; cc += kWarps) {
                const int jc = s.idx[cc];
                for (int a = cc + lane; a < nF; a += 32) Lb[cc * LD + a] = Qs[s.idx[a] * LD + jc];
            }
            for (int a = tid; a < nF; a += kThreads) {
                const int i = s.idx[a];
                T acc = T(0);
                for (int m = 0; m < nU; ++m) acc += sym_at(Qs, LD, i, s.idxU[m]);
                Lb[a * LD + nF] = s.r[i] - T(2) * u * acc;
                Lb[a * LD + nF + 1] = T(1);
            }
            __syncthreads();
            // (c) factor, eliminating the right-hand sides on the way
            lifted |= factor_ldl<T, kThreads>(Lb, LD, nF, 2, s.dinv, dmin);
        }
        // (d,e) multiplier of the budget constraint and back substitution -- warp 0
        if (warp == 0) {
            T s12 = T(0), s22 = T(0);
            for (int k = lane; k < nF; k += 32) {
                const T y1 = Lb[k * LD + nF], y2 = Lb[k * LD + nF + 1], di = s.dinv[k];
                s12 = fma(y1 * y2, di, s12);
                s22 = fma(y2 * y2, di, s22);
            }
            s12 = warp_sum(s12); s22 = warp_sum(s22);
            const T nu = (s12 - T(2) * c) / s22;
            T zt[NREG];
#pragma unroll
            for (int m = 0; m < NREG; ++m) {
                const int k = lane + 32 * m;
                zt[m] = k < nF ? T(0.5) * (Lb[k * LD + nF] - nu * Lb[k * LD + nF + 1]) : T(0);
            }
            backsub_warp<T, NREG>(Lb, LD, nF, s.dinv, zt);
#pragma unroll
            for (int m = 0; m < NREG; ++m) {
                const int k = lane + 32 * m;
                if (k < nF) s.wv[s.idx[k]] = zt[m];
            }
            if (lane == 0) s.red[32] = nu;
        }
        __syncthreads();
        nu_mult = s.red[32];
        // (f,g) gradient of the Lagrangian and the new active set
        for (int i = tid; i < N; i += kThreads) {
            T acc = T(0);
            for (int m = 0; m < nF; ++m) { const int j = s.idx[m]; acc = fma(sym_at(Qs, LD, i, j), s.wv[j], acc); }
            for (int m = 0; m < nU; ++m) acc = fma(sym_at(Qs, LD, i, s.idxU[m]), u, acc);
            const T gi = T(2) * acc - s.r[i] + nu_mult;
            const int8_t so = s.st[i];
            int8_t sn = so;
            if (so == 0) {
                const T wi = s.wv[i];
                if (wi < -tolw) sn = -1;
                else if (wi > u + tolw) sn = 1;
            } else if (so < 0) { if (gi < T(0)) sn = 0; }
            else { if (gi > T(0)) sn = 0; }
            if (sn != so) { s.st[i] = sn; changed = 1; }
        }
    } else {
        // no free asset: w is fixed by the caps; optimal iff max_L(-h) <= min_U(-h) and u |U| = 1
        for (int i = tid; i < N; i += kThreads) {
            T acc = T(0);
            for (int m = 0; m < nU; ++m) acc = fma(sym_at(Qs, LD, i, s.idxU[m]), u, acc);
            s.g[i] = T(2) * acc - s.r[i];
        }
        __syncthreads();
        if (warp == 0) {
            T hl = Num<T>::inf(), hu = Num<T>::inf();  // min h over L, min (-h) over U
            int il = 0x7fffffff, iu = 0x7fffffff;
            for (int i = lane; i < N; i += 32) {
                const T h = s.g[i];
                if (s.st[i] < 0) { if (h < hl) { hl = h; il = i; } }
                else if (s.st[i] > 0) { if (-h < hu) { hu = -h; iu = i; } }
            }
            warp_argmin(hl, il); warp_argmin(hu, iu);
            const T ctol = T(4) * Num<T>::eps() * (T)N;
            int ch = 0;
            if (lane == 0) {
                if (c > ctol && il < N) { s.st[il] = 0; ch = 1; }
                else if (c < -ctol && iu < N) { s.st[iu] = 0; ch = 1; }
                else if (il < N && iu < N && -hl > hu) { s.st[il] = 0; s.st[iu] = 0; ch = 1; }
                s.misc[2] = ch;
            }
        }
        __syncthreads();
        changed = s.misc[2];
    }
    return __syncthreads_or(changed);
}

// Safeguarded Newton on nu for the separable problem with diag(Q): initial active set.  Warp 0.
template <typename T, int NREG>
__device__ __forceinline__ void diag_init_warp(const T *qdiag, int qstride, const T *r, int8_t *st_out, int N, T u,
                                               T *dmax_out) {
    const int lane = threadIdx.x & 31;
    T i2q[NREG], rr[NREG];
    T lo = Num<T>::inf(), hi = -Num<T>::inf(), sa = T(0), sb = T(0), dmax = T(0);
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        const int i = lane + 32 * m;
        if (i < N) {
            T q = qdiag[i * qstride];
            dmax = q > dmax ? q : dmax;
            q = q > Num<T>::tiny() ? q : Num<T>::tiny();
            i2q[m] = T(0.5) / q; rr[m] = r[i];
            const T l = rr[m] - T(2) * q * u;
            lo = l < lo ? l : lo; hi = rr[m] > hi ? rr[m] : hi;
            sa = fma(rr[m], i2q[m], sa); sb += i2q[m];
        } else { i2q[m] = T(0); rr[m] = T(0); }
    }
    lo = warp_min(lo); hi = warp_max(hi); sa = warp_sum(sa); sb = warp_sum(sb); dmax = warp_max(dmax);
    T nu = (sa - T(1)) / sb;
    nu = clampT(nu, lo, hi);
    T dxold = hi - lo, dx = dxold;
    // keep this a loop (a single warp runs it; straight-line copies would be instruction-fetch bound)
#pragma unroll 1
    for (int it = 0; it < 60; ++it) {
        T sum = T(0), slope = T(0);
#pragma unroll
        for (int m = 0; m < NREG; ++m) {
            const T v = (rr[m] - nu) * i2q[m];
            sum += clampT(v, T(0), u);
            slope += (v > T(0) && v < u) ? i2q[m] : T(0);
        }
        sum = warp_sum(sum); slope = warp_sum(slope);
        // this is only the warm start of the active-set iteration: stop as soon as the budget holds to rounding
        // (a tolerance below the rounding of the sum itself made float runs bisect down to one ulp: measured 23k
        // cycles per portfolio)
        if (absT(sum - T(1)) <= T(128) * Num<T>::eps()) break;
        if (sum > T(1)) lo = nu; else hi = nu;
        T nn = nu;
        bool ok = false;
        if (slope > T(0)) {
            nn = nu + (sum - T(1)) / slope;
            ok = nn > lo && nn < hi && absT(T(2) * (nn - nu)) <= absT(dxold);
        }
        dxold = dx;
        if (ok) dx = nn - nu;
        else { dx = T(0.5) * (hi - lo); nn = lo + dx; }
        if (nn == nu) break;
        nu = nn;
    }
#pragma unroll
    for (int m = 0; m < NREG; ++m) {
        const int i = lane + 32 * m;
        if (i < N) {
            const T v = (rr[m] - nu) * i2q[m];
            st_out[i] = v <= T(0) ? -1 : (v >= u ? 1 : 0);
        }
    }
    *dmax_out = dmax;
}

// write the result of a converged / abandoned active-set iteration.  Returns (block-wide) whether a weight is not finite:
// NaN / Inf inputs make every comparison of the active-set step false, which reads as "converged" -- the caller turns the
// flag into DDW_STATUS_INEXACT so that the host layer raises SolverError instead of passing NaN weights on.
template <typename T, int kThreads>
__device__ __forceinline__ int write_result(const MkSmem<T> &s, int N, T u, T *w_out, int8_t *st_out) {
    int bad = 0;
    for (int i = threadIdx.x; i < N; i += kThreads) {
        int8_t st = s.st[i];
        const T wi = st == 0 ? clampT(s.wv[i], T(0), u) : (st > 0 ? u : T(0));
        // tie rule: a free asset whose weight lands exactly on a bound (within rounding) is reported as
        // active, so the saved state is a function of the returned weights
        if (st == 0) st = wi <= T(0) ? -1 : (wi >= u ? 1 : 0);
        bad |= !(absT(s.r[i]) < Num<T>::inf()) || (s.st[i] == 0 && !(absT(s.wv[i]) < Num<T>::inf()));   // NaN compares false
        w_out[i] = wi;
        st_out[i] = st;
    }
    return __syncthreads_or(bad);
}

// returns 1 when the problem was decided by the box alone (infeasible or a single feasible point)
template <typename T, int kThreads>
__device__ __forceinline__ int trivial_cases(int N, T u, T *w_out, int8_t *st_out, int32_t *status, int *bad_count) {
    const T cap = u * (T)N;
    const T tol = T(8) * Num<T>::eps() * (T)N;
    if (cap < T(1) - tol || !(u > T(0))) {
        for (int i = threadIdx.x; i < N; i += kThreads) { w_out[i] = Num<T>::inf() - Num<T>::inf(); st_out[i] = 0; }
        if (threadIdx.x == 0) { *status = DDW_STATUS_INFEASIBLE; atomicAdd(bad_count, 1); }
        return 1;
    }
    if (cap <= T(1) + tol) {
        for (int i = threadIdx.x; i < N; i += kThreads) { w_out[i] = u; st_out[i] = 1; }
        if (threadIdx.x == 0) *status = DDW_STATUS_OK;
        return 1;
    }
    return 0;
}

}  // namespace ddw
