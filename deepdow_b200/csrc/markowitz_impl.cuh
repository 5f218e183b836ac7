// markowitz_impl.cuh -- batched box-and-simplex QP solver and its implicit-differentiation backward.
//
// Replaces the CvxpyLayer call of NumericalMarkowitz.forward (reference
// deepdow/layers/allocate.py:240-273; problem stated at :220-232) and the diffcp adjoint behind it.
//
//   minimise  w'Qw - r'w   s.t. 1'w = 1, 0 <= w <= u,     Q = A'A + |alpha| I,
//   A[j,k] = gamma_sqrt[(b N^2 + j N + k) % B] * covmat_sqrt[b,j,k]      (the repeat/view tiling, :268-270)
//
// One CTA per portfolio.  Q (lower triangle) and the factor of the reduced KKT system share one
// N x LD shared-memory array: Q(i,j), j <= i, at Qs[i*LD + j]; the factor L(i,k), k <= i, of the
// free-set block plus two right-hand-side rows at Qs[k*LD + i + 1] (the strictly upper part).
//
// Forward = primal-dual active set (semismooth Newton on the KKT system) warm-started from the
// diagonal approximation of Q; every iterate is an exact reduced-KKT solve, so the converged
// answer is the exact optimum of the program that SCS only approximates.  Problems on which the
// active set cycles are handed to a Mehrotra interior-point kernel (same primitives) and polished.
#pragma once
#include "markowitz_common.cuh"


namespace ddw {

// sparse-support kernels (markowitz_sparse.cuh / markowitz_bwd_sparse.cuh, compiled in their own translation units)
bool fwd_sparse_eligible(int N, int tsize, int smem_limit);
template <typename T> int launch_fwd_sparse(const MkFwdParams<T> &p, cudaStream_t st);
bool bwd_sparse_eligible(int N, int tsize, int smem_limit);
template <typename T> int launch_bwd_sparse(const MkBwdParams<T> &p, cudaStream_t st);

// markowitz_qwarp.cu: float32, n_assets <= 128 -- Q on tensor cores into an L2-resident scratch, then one warp per portfolio
bool fwd_qwarp_eligible(int N, int tsize);
size_t fwd_qwarp_scratch_bytes(long long B, int N);
int launch_fwd_qwarp(const MkFwdParams<float> &p, float *Q, cudaStream_t st);

// markowitz_gram_tc.cu: Q = A'A + |alpha| I on tensor cores into the global workspace (float32, Q beyond shared memory)
int markowitz_gram_tc(const float *C, const float *gamma, const float *alpha, long long B, long long Bmod, long long b0, int N,
                      int LD, float *Q, cudaStream_t st);

// Dense forward kernel.  Direct mode (work_list == nullptr): one CTA per portfolio, grid = B.
// List mode: a fixed grid walks the portfolios that the sparse kernel handed over.
// MAXT > 0: Gram tiles in registers; 0: Gram through the shared staging area; < 0: Q was formed by markowitz_gram_tc.
// kFused: covmat_sqrt = CovarianceMatrix(sqrt=False)(X) is formed in shared memory from the (L, N) lookback tile (cov_tile.cuh,
// the code of covariance_plain_kernel: same bits) instead of being read from HBM; direct mode only, 128 threads.
template <typename T, int kThreads, int NREG, bool kGlobal, int MAXT, bool kFused = false>
__global__ void __launch_bounds__(kThreads) markowitz_fwd_kernel(MkFwdParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.N, LD = p.LD, tid = threadIdx.x;
    const long long nwork = p.work_list ? (long long)*p.work_count : p.B;
    for (long long wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
        const long long b = p.work_list ? (long long)p.work_list[wi] : wi;
        T *w_out = p.w + b * N;
        int8_t *st_out = p.state + b * N;
        if (trivial_cases<T, kThreads>(N, p.u, w_out, st_out, p.status + b, p.bad_count)) continue;
        MkSmem<T> s = carve<T, kGlobal>(smem_raw, N, LD, p.AS, p.CH, p.ws, b);
        __syncthreads();   // list mode: the previous portfolio's shared state is dead from here on
        for (int i = tid; i < N; i += kThreads) s.r[i] = p.rets[b * N + i];
        const T aabs = absT(p.alpha[b]);
        const T *Cb = p.C + b * (long long)N * N;
        if (kFused) {
            // the covariance tile lives behind the solver's own shared memory
            const int XS = round_up(N, 4), L = p.L;
            T *X = reinterpret_cast<T *>(smem_raw + round_up((int)mk_smem_bytes(N, LD, p.AS, p.CH, sizeof(T), false), 16));
            T *Xc = X + round_up(L * N, 4);
            T *S = Xc + L * XS;
            T *part = S + round_up(N * N, 4);
            T *red2 = part + kPlainParts * XS;
            const T *xb = p.X + b * (long long)L * N;
            for (int e = tid; e < L * N; e += kThreads) X[e] = xb[e];
            __syncthreads();
            cov_plain_tile<T>(X, Xc, S, part, red2, L, N, XS, p.strategy, p.strategy == DDW_SHRINK_NONE ? T(1) : p.coef[b]);
            Cb = S;
        }
        if (MAXT > 0)
            gram_lower_reg<T, kThreads, (MAXT > 0 ? MAXT : 1), false>(Cb, p.gamma,
                                                                       (p.b0 + b) * (long long)N * N, p.Bmod, N, N, nullptr, nullptr,
                                                                       N, s.As, p.AS, p.CH, s.Qs, LD, aabs);
        else if (MAXT == 0)
            gram_lower<T, kThreads, false>(Cb, p.gamma, (p.b0 + b) * (long long)N * N, p.Bmod, N, N,
                                           nullptr, nullptr, N, s.As, p.AS, p.CH, s.Qs, LD, aabs);
        else
            __syncthreads();          // rets staged; Q already sits in the global workspace
        if (tid < 32) {
            T dmax;
            diag_init_warp<T, NREG>(s.Qs, LD + 1, s.r, s.st, N, p.u, &dmax);
            if (tid == 0) s.red[33] = dmax;
        }
        __syncthreads();
        const T dmin = T(4) * Num<T>::eps() * s.red[33] + Num<T>::tiny();
        const T tolw = T(8) * Num<T>::eps();
        int lifted = 0, converged = 0;
        for (int it = 0; it < kMaxPdas; ++it) {
            if (!pdas_step<T, kThreads, NREG>(s, N, LD, p.u, dmin, tolw, lifted)) { converged = 1; break; }
        }
        const int nonfinite = write_result<T, kThreads>(s, N, p.u, w_out, st_out);
        if (p.saved_n) {
            // keep the factor of the converged reduced KKT system for the backward: the last active-set step factored exactly
            // the final free set.  Not when a free weight landed on a bound (the reported state then differs from that set).
            const int nF = s.misc[0];
            int tie = 0;
            for (int i = tid; i < N; i += kThreads)
                if (s.st[i] == 0) { const T wi = s.wv[i]; tie |= (wi <= T(0)) || (wi >= p.u); }
            tie = __syncthreads_or(tie);
            const bool keep = converged && !nonfinite && !tie;
            if (keep && nF > 0) {
                const T *Lb = s.Qs + 1;
                if (kGlobal) {
                    for (int k = tid; k < nF; k += kThreads) s.Qs[(long long)k * LD + k] = s.dinv[k];   // Q is dead after convergence
                } else {
                    T *dst = p.saved + b * p.saved_stride;
                    for (int cc = tid >> 5; cc < nF; cc += kThreads / 32) {
                        const int off = cc * nF - cc * (cc - 1) / 2 - cc;        // column cc holds rows cc .. nF-1
                        for (int a = cc + (tid & 31); a < nF; a += 32) dst[off + a] = Lb[cc * LD + a];
                    }
                    for (int k = tid; k < nF; k += kThreads) dst[nF * (nF + 1) / 2 + k] = s.dinv[k];
                }
            }
            if (tid == 0) p.saved_n[b] = keep ? nF : -1;
        }
        if (kFused && !converged && !nonfinite && p.C) {
            // the interior-point fallback reads covmat_sqrt from HBM: spill this portfolio's covariance tile for it
            T *dst = const_cast<T *>(p.C) + b * (long long)N * N;
            for (int e = tid; e < N * N; e += kThreads) dst[e] = Cb[e];
        }
        if (tid == 0) {
            int st = lifted ? DDW_STATUS_REGULARIZED : 0;
            if (nonfinite) { st |= DDW_STATUS_INEXACT; atomicAdd(p.bad_count, 1); }     // NaN / Inf in the inputs
            else if (!converged) {
                st |= DDW_STATUS_FALLBACK | DDW_STATUS_INEXACT;
                if (kFused && !p.C) atomicAdd(p.bad_count, 1);      // no covmat_sqrt in HBM for the fallback kernel
                else {
                    const int slot = atomicAdd(p.fail_count, 1);
                    p.fail_list[slot] = (int)b;
                }
            }
            p.status[b] = st;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fallback: Mehrotra predictor-corrector interior point on the same data, then active-set polish.
// Launched with a fixed small grid; CTAs walk the list of failed problems.
// ------------------------------------------------------------------------------------------------
template <typename T, int kThreads, int NREG, bool kGlobal>
__global__ void __launch_bounds__(kThreads) markowitz_ipm_kernel(MkFwdParams<T> p, T *ipm_ws, int have_q) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.N, LD = p.LD, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NP = round_up(N + 2, 32);
    const int nfail = *p.fail_count;
    for (int f = blockIdx.x; f < nfail; f += gridDim.x) {
        const long long b = p.fail_list[f];
        MkSmem<T> s = carve<T, kGlobal>(smem_raw, N, LD, p.AS, p.CH, p.ws, b);
        // per-CTA vectors of the interior point method live in global scratch (rare path)
        T *wk = ipm_ws + (long long)blockIdx.x * 8 * NP;
        T *w = wk, *lam = wk + NP, *mu = wk + 2 * NP, *h = wk + 3 * NP, *dw = wk + 4 * NP, *dl = wk + 5 * NP,
          *dm = wk + 6 * NP, *y2 = wk + 7 * NP;
        const T u = p.u;
        __syncthreads();
        for (int i = tid; i < N; i += kThreads) s.r[i] = p.rets[b * N + i];
        const T aabs = absT(p.alpha[b]);
        // have_q: the lower triangle of Q in the global workspace is still intact (the active-set kernel only wrote the factor
        // into the strictly upper part)
        if (!(kGlobal && have_q))
            gram_lower<T, kThreads, false>(p.C + b * (long long)N * N, p.gamma, (p.b0 + b) * (long long)N * N, p.Bmod, N, N, nullptr,
                                           nullptr, N, s.As, p.AS, p.CH, s.Qs, LD, aabs);
        else
            __syncthreads();
        // objective scale so that gradients are O(1)
        T sc = T(0);
        for (int i = tid; i < N; i += kThreads) {
            const T a = absT(s.r[i]), q2 = T(2) * s.Qs[i * LD + i];
            sc = a > sc ? a : sc; sc = q2 > sc ? q2 : sc;
        }
        sc = warp_max(sc);
        if (lane == 0) s.red[warp] = sc;
        __syncthreads();
        sc = T(0);
        for (int i = 0; i < kThreads / 32; ++i) sc = s.red[i] > sc ? s.red[i] : sc;
        __syncthreads();
        if (!(sc > T(0))) sc = T(1);
        const T isc = T(1) / sc;
        const bool useU = u < T(1);
        const T ncomp = (T)(N * (useU ? 2 : 1));
        for (int i = tid; i < N; i += kThreads) {
            w[i] = T(1) / (T)N;
            lam[i] = (T)N;
            mu[i] = useU ? T(1) / (u - w[i]) : T(0);
        }
        T nu = T(0);
        T *Qs = s.Qs, *Lb = s.Qs + 1;
        const T gtol = T(64) * Num<T>::eps();
        __syncthreads();
        for (int it = 0; it < kMaxIpm; ++it) {
            // residuals: h = (2 Q w - r)/sc + nu
            T lgap = T(0), lrp = T(0);
            for (int i = tid; i < N; i += kThreads) {
                T acc = T(0);
                for (int j = 0; j < N; ++j) acc = fma(sym_at(Qs, LD, i, j), w[j], acc);
                h[i] = (T(2) * acc - s.r[i]) * isc + nu;
                lgap += lam[i] * w[i] + (useU ? mu[i] * (u - w[i]) : T(0));
                lrp += w[i];
            }
            const T gap = block_sum<T, kThreads>(lgap, s.red) / ncomp;
            const T rp = block_sum<T, kThreads>(lrp, s.red) - T(1);
            if (gap < gtol) break;
            // H = 2Q/sc + diag(lam/w + mu/t); rhs rows: [1, -h]
            for (int cc = warp; cc < N; cc += kThreads / 32)
                for (int a = cc + lane; a < N; a += 32) {
                    T v = T(2) * Qs[a * LD + cc] * isc;
                    if (a == cc) v += lam[a] / w[a] + (useU ? mu[a] / (u - w[a]) : T(0));
                    Lb[cc * LD + a] = v;
                }
            for (int a = tid; a < N; a += kThreads) { Lb[a * LD + N] = T(1); Lb[a * LD + N + 1] = -h[a]; }
            __syncthreads();
            factor_ldl<T, kThreads>(Lb, LD, N, 2, s.dinv, Num<T>::tiny());
            T s2 = T(0), s1 = T(0);
            if (warp == 0) {
                T z1[NREG], z2[NREG];
#pragma unroll
                for (int m = 0; m < NREG; ++m) {
                    const int k = lane + 32 * m;
                    z1[m] = k < N ? Lb[k * LD + N] : T(0);
                    z2[m] = k < N ? Lb[k * LD + N + 1] : T(0);
                }
                backsub_warp<T, NREG>(Lb, LD, N, s.dinv, z1);
                backsub_warp<T, NREG>(Lb, LD, N, s.dinv, z2);
#pragma unroll
                for (int m = 0; m < NREG; ++m) {
                    const int k = lane + 32 * m;
                    if (k < N) { y2[k] = z1[m]; dw[k] = z2[m]; s2 += z1[m]; s1 += z2[m]; }
                }
                s2 = warp_sum(s2); s1 = warp_sum(s1);
                if (lane == 0) { s.red[34] = s2; s.red[35] = s1; }
            }
            __syncthreads();
            s2 = s.red[34]; s1 = s.red[35];
            const T dnu = (s1 + rp) / s2;
            // predictor step and its maximal step length
            T amax = T(1);
            for (int i = tid; i < N; i += kThreads) {
                const T d = dw[i] - dnu * y2[i];
                dw[i] = d;
                const T dli = -lam[i] - lam[i] / w[i] * d;
                dl[i] = dli;
                if (d < T(0)) amax = min(amax, -w[i] / d);
                if (dli < T(0)) amax = min(amax, -lam[i] / dli);
                if (useU) {
                    const T t = u - w[i];
                    const T dmi = -mu[i] + mu[i] / t * d;
                    dm[i] = dmi;
                    if (d > T(0)) amax = min(amax, t / d);
                    if (dmi < T(0)) amax = min(amax, -mu[i] / dmi);
                } else dm[i] = T(0);
            }
            amax = warp_min(amax);
            if (lane == 0) s.red[warp] = amax;
            __syncthreads();
            amax = T(1);
            for (int i = 0; i < kThreads / 32; ++i) amax = min(amax, s.red[i]);
            __syncthreads();
            T lga = T(0);
            for (int i = tid; i < N; i += kThreads)
                lga += (lam[i] + amax * dl[i]) * (w[i] + amax * dw[i]) +
                       (useU ? (mu[i] + amax * dm[i]) * (u - w[i] - amax * dw[i]) : T(0));
            const T gap_aff = block_sum<T, kThreads>(lga, s.red) / ncomp;
            T sigma = gap_aff / gap;
            sigma = sigma * sigma * sigma;
            // corrector right-hand side, forward + backward substitution with the same factor
            for (int i = tid; i < N; i += kThreads) {
                const T cl = (sigma * gap - dw[i] * dl[i]) / w[i];
                const T cu = useU ? (sigma * gap + dw[i] * dm[i]) / (u - w[i]) : T(0);
                s.g[i] = -h[i] + cl - cu;   // rhs
                s.wv[i] = cl;               // cu is recomputed below
            }
            __syncthreads();
            if (warp == 0) {
                T z[NREG];
#pragma unroll
                for (int m = 0; m < NREG; ++m) { const int k = lane + 32 * m; z[m] = k < N ? s.g[k] : T(0); }
                fwdsub_warp<T, NREG>(Lb, LD, N, s.dinv, z);
                backsub_warp<T, NREG>(Lb, LD, N, s.dinv, z);
                T ss = T(0);
#pragma unroll
                for (int m = 0; m < NREG; ++m) { const int k = lane + 32 * m; if (k < N) { s.g[k] = z[m]; ss += z[m]; } }
                ss = warp_sum(ss);
                if (lane == 0) s.red[36] = ss;
            }
            __syncthreads();
            const T dnu2 = (s.red[36] + rp) / s2;
            amax = T(1);
            for (int i = tid; i < N; i += kThreads) {
                const T d2 = s.g[i] - dnu2 * y2[i];
                const T cl = s.wv[i];
                const T cu = useU ? (sigma * gap + dw[i] * dm[i]) / (u - w[i]) : T(0);
                const T dl2 = -lam[i] + cl - lam[i] / w[i] * d2;
                s.g[i] = d2; dl[i] = dl2;
                if (d2 < T(0)) amax = min(amax, -w[i] / d2);
                if (dl2 < T(0)) amax = min(amax, -lam[i] / dl2);
                if (useU) {
                    const T t = u - w[i];
                    const T dm2 = -mu[i] + cu + mu[i] / t * d2;
                    dm[i] = dm2;
                    if (d2 > T(0)) amax = min(amax, t / d2);
                    if (dm2 < T(0)) amax = min(amax, -mu[i] / dm2);
                }
            }
            amax = warp_min(amax);
            if (lane == 0) s.red[warp] = amax;
            __syncthreads();
            amax = T(1);
            for (int i = 0; i < kThreads / 32; ++i) amax = min(amax, s.red[i]);
            __syncthreads();
            const T a = min(T(1), T(0.99) * amax);
            for (int i = tid; i < N; i += kThreads) {
                w[i] += a * s.g[i];
                lam[i] += a * dl[i];
                if (useU) mu[i] += a * dm[i];
            }
            nu += a * dnu2;
            __syncthreads();
        }
        // active-set identification + exact polish
        for (int i = tid; i < N; i += kThreads) {
            int8_t st = 0;
            if (w[i] < lam[i]) st = -1;
            if (useU && (u - w[i]) < mu[i]) st = 1;
            s.st[i] = st;
            y2[i] = w[i];  // keep the interior-point iterate
        }
        if (tid < 32) {
            T dmax = T(0);
            for (int i = lane; i < N; i += 32) dmax = max(dmax, Qs[i * LD + i]);
            dmax = warp_max(dmax);
            if (lane == 0) s.red[37] = dmax;
        }
        __syncthreads();
        const T dmin = T(4) * Num<T>::eps() * s.red[37] + Num<T>::tiny();
        int lifted = 0, converged = 0;
        for (int it = 0; it < kMaxPdas; ++it) {
            if (!pdas_step<T, kThreads, NREG>(s, N, LD, u, dmin, T(8) * Num<T>::eps(), lifted)) { converged = 1; break; }
        }
        if (converged) {
            const int nonfinite = write_result<T, kThreads>(s, N, u, p.w + b * N, p.state + b * N);
            if (tid == 0) {
                p.status[b] = DDW_STATUS_FALLBACK | (lifted ? DDW_STATUS_REGULARIZED : 0) | (nonfinite ? DDW_STATUS_INEXACT : 0);
                if (nonfinite) atomicAdd(p.bad_count, 1);
            }
        } else {
            const T thr = T(1e3) * Num<T>::eps();
            for (int i = tid; i < N; i += kThreads) {
                const T wi = clampT(y2[i], T(0), u);
                p.w[b * N + i] = wi;
                p.state[b * N + i] = wi <= thr ? -1 : (u - wi <= thr ? 1 : 0);
            }
            if (tid == 0) { p.status[b] = DDW_STATUS_FALLBACK | DDW_STATUS_INEXACT; atomicAdd(p.bad_count, 1); }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Backward: implicit differentiation of the KKT system on the active set saved by the forward.
//   F = {state == 0};  [[2 Q_FF, 1],[1', 0]] [d_F; eta] = [g_F; 0],  d = 0 off F
//   d_rets = d;  dA = -2 [(A d) w' + (A w) d'];  d_covmat_sqrt = gamma_ * dA;  d_alpha = sign(alpha)(-2 d.w)
// The (A d, A w) vectors are also written to `vecs` (B,2,N) for the gamma scatter pass.
// ------------------------------------------------------------------------------------------------
static size_t mk_bwd_smem_bytes(int N, int LDW, int AS, int CH, int tsize, bool global_l) {
    const int NP = round_up(N + 2, 32);
    size_t elems = (global_l ? 0 : (size_t)round_up(N * LDW, 4)) + (size_t)2 * CH * AS + (size_t)7 * NP + 64;
    return elems * tsize + ((size_t)3 * NP + 8) * sizeof(int);
}

// Two paths.  "Compact" (the usual case): the columns of A that carry weight (free + capped assets)
// are copied once from HBM into shared memory, and the restricted Gram, A d and A w all read that
// copy.  "Streaming" (support too large for the arena, or Q in global memory): chunked Gram over the
// free columns and a second, L2-resident pass over A for A d and A w.
template <typename T, int kThreads, int NREG, bool kGlobal>
__global__ void __launch_bounds__(kThreads) markowitz_bwd_kernel(MkBwdParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // list mode: the sparse-support kernel handed over the portfolios whose free set is too large for it
    if (p.work_list && (long long)blockIdx.x >= (long long)*p.work_count) return;
    const long long b = p.work_list ? (long long)p.work_list[blockIdx.x] : (long long)blockIdx.x;
    const int N = p.N, LDW = p.LDW, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;
    const int NP = round_up(N + 2, 32);
    T *ptr = reinterpret_cast<T *>(smem_raw);
    T *arena = ptr;
    const int arena_elems = (kGlobal ? 0 : round_up(N * LDW, 4)) + 2 * p.CH * p.AS;
    ptr += arena_elems;
    T *wv = ptr; ptr += NP;
    T *dv = ptr; ptr += NP;
    T *gv = ptr; ptr += NP;
    T *dinv = ptr; ptr += NP;
    T *qv = ptr; ptr += NP;
    T *wc = ptr; ptr += NP;     // w and d on the compacted support
    T *dc = ptr; ptr += NP;
    T *red = ptr; ptr += 64;
    int *idx = reinterpret_cast<int *>(ptr);
    int *pos = idx + NP;        // column -> index among the free assets, or -1
    int *posS = pos + NP;       // column -> index among free + capped assets, or -1
    int *misc = posS + NP;

    if (warp == 0) {
        int nf = 0, nu = 0;
        for (int base = 0; base < N; base += 32) {
            const int i = base + lane;
            const int v = i < N ? (int)p.state[b * N + i] : 2;
            const unsigned mf = __ballot_sync(kFullMask, v == 0);
            const int a = nf + __popc(mf & ((1u << lane) - 1u));
            if (i < N) pos[i] = v == 0 ? a : -1;
            if (v == 0) idx[a] = i;
            nf += __popc(mf);
        }
        for (int base = 0; base < N; base += 32) {
            const int i = base + lane;
            const int v = i < N ? (int)p.state[b * N + i] : 2;
            const unsigned mu = __ballot_sync(kFullMask, v == 1);
            if (i < N) posS[i] = v == 0 ? pos[i] : (v == 1 ? nf + nu + __popc(mu & ((1u << lane) - 1u)) : -1);
            nu += __popc(mu);
        }
        if (lane == 0) { misc[0] = nf; misc[1] = nu; }
    }
    for (int i = tid; i < N; i += kThreads) {
        wv[i] = p.w[b * N + i];
        gv[i] = p.grad_w[b * N + i];
        dv[i] = T(0);
    }
    for (int i = tid; i < NP; i += kThreads) { wc[i] = T(0); dc[i] = T(0); }
    __syncthreads();
    const int nF = misc[0], nS = misc[0] + misc[1];
    const T *Cb = p.C + b * (long long)N * N;
    const long long gflat0 = (p.b0 + b) * (long long)N * N;
    const T alpha = p.alpha[b];
    const unsigned uB = (unsigned)p.Bmod;
    const int CS = round_up(nS, 4), LDc = (nF + 2) | 1;
    const int NT = (nF + 3) >> 2, ntiles = NT * (NT + 1) / 2;
    // K-split of the small Gram: KS slices of rows per tile, partial sums reduced through shared memory
    int KS = 1;
    if (ntiles > 0) { KS = kThreads / ntiles; KS = KS < 1 ? 1 : (KS > 8 ? 8 : KS); }
    const bool compact = !kGlobal && nF > 0 &&
                         (N * CS + round_up(nF * LDc, 4) + (KS > 1 ? KS * ntiles * 16 : 0) <= arena_elems);
    T *pv = gv;   // p = A d overwrites grad_w once d is known
    T *Ac = arena, *Lb, *part = nullptr;
    int ldl;
    if (compact) { Lb = arena + N * CS; ldl = LDc; part = Lb + round_up(nF * LDc, 4); }
    else { Lb = kGlobal ? p.ws + b * (long long)N * LDW : arena; ldl = LDW; }

    if (nF > 0) {
        if (compact) {
            // ---- one pass over A: keep the support columns (gamma-tiled) in shared memory
            for (int i = tid; i < N; i += kThreads) { const int c = posS[i]; if (c >= 0) wc[c] = wv[i]; }
            if ((N & 3) == 0) {
                const int n4 = N * N / 4;
                int k = (4 * tid) / N, i = 4 * tid - k * N;
                const int dk = (4 * kThreads) / N, di = (4 * kThreads) % N;
                unsigned gi = (unsigned)((gflat0 + 4 * tid) % p.Bmod);
                const unsigned gstep = (unsigned)((4 * kThreads) % p.Bmod);
                const bool gvec = (uB & 3u) == 0u;   // then every aligned 4-group of flat indices is contiguous mod B
                for (int e4 = tid; e4 < n4; e4 += 4 * kThreads) {
                    Vec4<T> v[4], g4[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (e4 + r * kThreads < n4) v[r] = ld4(Cb + 4 * (e4 + r * kThreads));
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (e4 + r * kThreads < n4) {
                            if (gvec) g4[r] = ld4(p.gamma + gi);
                            else { unsigned g = gi;
#pragma unroll
                                for (int q = 0; q < 4; ++q) { g4[r].v[q] = p.gamma[g]; if (++g >= uB) g = 0; } }
                        }
                        gi += gstep; if (gi >= uB) gi -= uB;
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (e4 + r * kThreads < n4) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int c = posS[i + q];
                                if (c >= 0) Ac[k * CS + c] = v[r].v[q] * g4[r].v[q];
                            }
                        }
                        i += di; k += dk;
                        if (i >= N) { i -= N; ++k; }
                    }
                }
            } else {
                int k = tid / N, i = tid - k * N;
                const int dk = kThreads / N, di = kThreads % N;
                unsigned gi = (unsigned)((gflat0 + tid) % p.Bmod);
                const unsigned gstep = (unsigned)(kThreads % p.Bmod);
                for (int e = tid; e < N * N; e += kThreads) {
                    const int c = posS[i];
                    if (c >= 0) Ac[k * CS + c] = Cb[e] * p.gamma[gi];
                    i += di; k += dk;
                    if (i >= N) { i -= N; ++k; }
                    gi += gstep; if (gi >= uB) gi -= uB;
                }
            }
            // pad columns of the compact copy are read by the 4-wide tile loads: keep them finite
            for (int e = tid; e < N * (CS - nS); e += kThreads) {
                const int k = e / (CS - nS), c = nS + e - k * (CS - nS);
                Ac[k * CS + c] = T(0);
            }
            __syncthreads();
            // ---- Q_FF = A_F' A_F + |alpha| I, 4x4 tiles, rows split KS ways
            {
                const int q = tid % KS, tstep = kThreads / KS;
                for (int t = tid / KS; t < ntiles; t += tstep) {
                    int ti = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
                    while (ti * (ti + 1) / 2 > t) --ti;
                    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
                    const int tj = t - ti * (ti + 1) / 2;
                    T acc[4][4];
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) acc[x][y] = T(0);
                    for (int k = q; k < N; k += KS) {
                        const Vec4<T> va = ld4(Ac + k * CS + 4 * ti), vb = ld4(Ac + k * CS + 4 * tj);
#pragma unroll
                        for (int x = 0; x < 4; ++x)
#pragma unroll
                            for (int y = 0; y < 4; ++y) acc[x][y] = fma(va.v[x], vb.v[y], acc[x][y]);
                    }
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
                            if (KS > 1) part[(q * ntiles + t) * 16 + 4 * x + y] = acc[x][y];
                            else {
                                const int i = 4 * ti + x, j = 4 * tj + y;
                                if (i < nF && j <= i) Lb[j * ldl + i] = acc[x][y] + (i == j ? absT(alpha) : T(0));
                            }
                        }
                }
                if (KS > 1) {
                    __syncthreads();
                    for (int e = tid; e < ntiles * 16; e += kThreads) {
                        const int tt = e >> 4, xy = e & 15;
                        int t2 = (int)((sqrtf(8.f * (float)tt + 1.f) - 1.f) * 0.5f);
                        while (t2 * (t2 + 1) / 2 > tt) --t2;
                        while ((t2 + 1) * (t2 + 2) / 2 <= tt) ++t2;
                        const int i = 4 * t2 + (xy >> 2), j = 4 * (tt - t2 * (t2 + 1) / 2) + (xy & 3);
                        if (i < nF && j <= i) {
                            T sum = T(0);
                            for (int qq = 0; qq < KS; ++qq) sum += part[(qq * ntiles + tt) * 16 + xy];
                            Lb[j * ldl + i] = sum + (i == j ? absT(alpha) : T(0));
                        }
                    }
                }
                __syncthreads();
            }
        } else {
            gram_lower<T, kThreads, true>(Cb, p.gamma, gflat0, p.Bmod, N, N, nullptr, pos, nF, arena + (kGlobal ? 0 : round_up(N * LDW, 4)),
                                          p.AS, p.CH, Lb, ldl, absT(alpha));
        }
        T dmax = T(0);
        for (int a = tid; a < nF; a += kThreads) {
            Lb[a * ldl + nF] = gv[idx[a]];
            Lb[a * ldl + nF + 1] = T(1);
            dmax = max(dmax, Lb[a * ldl + a]);
        }
        dmax = warp_max(dmax);
        if (lane == 0) red[warp] = dmax;
        __syncthreads();
        dmax = T(0);
        for (int i = 0; i < kWarps; ++i) dmax = max(dmax, red[i]);
        const T dmin = T(4) * Num<T>::eps() * dmax + Num<T>::tiny();
        if (small_system_fits<T, kThreads>(nF))
            factor_small<T, kThreads>(Lb, ldl, nF, dinv, dmin, [&](int i, int j) -> T { return Lb[j * ldl + i]; });
        else
            factor_ldl<T, kThreads>(Lb, ldl, nF, 2, dinv, dmin);
        if (warp == 0) {
            T s12 = T(0), s22 = T(0);
            for (int k = lane; k < nF; k += 32) {
                const T y1 = Lb[k * ldl + nF], y2 = Lb[k * ldl + nF + 1], di = dinv[k];
                s12 = fma(y1 * y2, di, s12);
                s22 = fma(y2 * y2, di, s22);
            }
            s12 = warp_sum(s12); s22 = warp_sum(s22);
            const T eta = s12 / s22;
            T zt[NREG];
#pragma unroll
            for (int m = 0; m < NREG; ++m) {
                const int k = lane + 32 * m;
                zt[m] = k < nF ? T(0.5) * (Lb[k * ldl + nF] - eta * Lb[k * ldl + nF + 1]) : T(0);
            }
            backsub_warp<T, NREG>(Lb, ldl, nF, dinv, zt);
#pragma unroll
            for (int m = 0; m < NREG; ++m) {
                const int k = lane + 32 * m;
                if (k < nF) { dv[idx[k]] = zt[m]; dc[k] = zt[m]; }
            }
        }
        __syncthreads();
    }
    // d_rets, d_alpha
    T ldot = T(0);
    for (int i = tid; i < N; i += kThreads) { p.d_rets[b * N + i] = dv[i]; ldot = fma(dv[i], wv[i], ldot); }
    const T dot = block_sum<T, kThreads>(ldot, red);
    if (tid == 0) p.d_alpha[b] = (alpha > T(0) ? T(1) : (alpha < T(0) ? T(-1) : T(0))) * (T(-2) * dot);
    // p = A d, q = A w
    if (compact) {
        for (int k = tid; k < N; k += kThreads) {
            T sp = T(0), sq = T(0);
            for (int c = 0; c < CS; c += 4) {
                const Vec4<T> a = ld4(Ac + k * CS + c), d4 = ld4(dc + c), w4 = ld4(wc + c);
#pragma unroll
                for (int q = 0; q < 4; ++q) { sp = fma(a.v[q], d4.v[q], sp); sq = fma(a.v[q], w4.v[q], sq); }
            }
            pv[k] = sp; qv[k] = sq;
        }
    } else if (nF > 0) {
        // one warp per row, coalesced reads of covmat_sqrt (L2 resident after the Gram)
        unsigned grow = (unsigned)((gflat0 + (long long)warp * N + lane) % p.Bmod);
        const unsigned growstep = (unsigned)(((long long)kWarps * N) % p.Bmod);
        const unsigned gstep = (unsigned)(32 % p.Bmod);
        for (int j = warp; j < N; j += kWarps) {
            T sp = T(0), sq = T(0);
            unsigned gi = grow;
            grow += growstep; if (grow >= uB) grow -= uB;
            for (int k = lane; k < N; k += 32) {
                const T a = Cb[(long long)j * N + k] * p.gamma[gi];
                sp = fma(a, dv[k], sp); sq = fma(a, wv[k], sq);
                gi += gstep; if (gi >= uB) gi -= uB;
            }
            sp = warp_sum(sp); sq = warp_sum(sq);
            if (lane == 0) { pv[j] = sp; qv[j] = sq; }
        }
    } else {
        for (int j = tid; j < N; j += kThreads) { pv[j] = T(0); qv[j] = T(0); }   // d = 0: every gradient vanishes
    }
    __syncthreads();
    if (p.vecs) for (int i = tid; i < N; i += kThreads) { p.vecs[(b * 2) * N + i] = pv[i]; p.vecs[(b * 2 + 1) * N + i] = qv[i]; }
    // d_covmat_sqrt = gamma_ * (-2) (p w' + q d'): one warp per row, 16-byte stores when rows are aligned
    T *dC = p.d_C + b * (long long)N * N;
    if ((N & 3) == 0) {
        // gamma index of (row j, column 4*lane): one 64-bit modulo per thread, then incremental
        unsigned grow = (unsigned)((gflat0 + (long long)warp * N + 4 * lane) % p.Bmod);
        const unsigned growstep = (unsigned)(((long long)kWarps * N) % p.Bmod);
        const unsigned gstep = (unsigned)(128 % p.Bmod);
        const bool gvec = (uB & 3u) == 0u;
        for (int j = warp; j < N; j += kWarps) {
            const T pj = T(-2) * pv[j], qj = T(-2) * qv[j];
            unsigned gi = grow;
            for (int k = 4 * lane; k < N; k += 128) {
                const Vec4<T> w4 = ld4(wv + k), d4 = ld4(dv + k);
                Vec4<T> o, g4;
                if (gvec) g4 = ld4(p.gamma + gi);
                else { unsigned g = gi;
#pragma unroll
                    for (int q = 0; q < 4; ++q) { g4.v[q] = p.gamma[g]; if (++g >= uB) g = 0; } }
#pragma unroll
                for (int q = 0; q < 4; ++q) o.v[q] = g4.v[q] * fma(pj, w4.v[q], qj * d4.v[q]);
                st4(dC + (long long)j * N + k, o);
                gi += gstep; if (gi >= uB) gi -= uB;
            }
            grow += growstep; if (grow >= uB) grow -= uB;
        }
    } else {
        const int total = N * N;
        int j = tid / N, k = tid - j * N;
        const int dj = kThreads / N, dkk = kThreads % N;
        unsigned gi = (unsigned)((gflat0 + tid) % p.Bmod);
        const unsigned gstep = (unsigned)(kThreads % p.Bmod);
        for (int e = tid; e < total; e += kThreads) {
            dC[e] = p.gamma[gi] * T(-2) * fma(pv[j], wv[k], qv[j] * dv[k]);
            k += dkk; j += dj;
            if (k >= N) { k -= N; ++j; }
            gi += gstep; if (gi >= uB) gi -= uB;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward from the factor the forward kept (ddw_markowitz_fwd_saved): no Gram, no factorisation -- two forward
// substitutions (g_F and 1), the budget multiplier, one back substitution, then the same streaming tail as above
// (A d, A w from covmat_sqrt, the rank-2 gradient).  kGlobal == false: the packed factor is unpacked into shared memory;
// true: the factor is read in place from the forward's global Q / L array (row stride LDF, 1 / pivot on Q's diagonal).
// Problems without a saved factor are appended to work_list2 for markowitz_bwd_kernel.
// ------------------------------------------------------------------------------------------------
static size_t mk_bwd_saved_smem_bytes(int N, int tsize, bool global_l) {
    const int NP = round_up(N + 2, 32);
    size_t elems = (global_l ? 0 : (size_t)round_up(N * (N | 1), 4)) + (size_t)6 * NP + 64;
    return elems * tsize + ((size_t)NP + 8) * sizeof(int);
}

template <typename T, int kThreads, int NREG, bool kGlobal>
__global__ void __launch_bounds__(kThreads) markowitz_bwd_saved_kernel(MkBwdParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (p.work_list && (long long)blockIdx.x >= (long long)*p.work_count) return;
    const long long b = p.work_list ? (long long)p.work_list[blockIdx.x] : (long long)blockIdx.x;
    const int N = p.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;
    const int nF = p.saved_n[b];
    if (nF < 0) {
        if (tid == 0) { const int sl = atomicAdd(p.work_count2, 1); p.work_list2[sl] = (int)b; }
        return;
    }
    const int NP = round_up(N + 2, 32);
    T *ptr = reinterpret_cast<T *>(smem_raw);
    T *Lsm = ptr; ptr += kGlobal ? 0 : round_up(N * (N | 1), 4);
    T *wv = ptr; ptr += NP;
    T *dv = ptr; ptr += NP;
    T *gv = ptr; ptr += NP;
    T *qv = ptr; ptr += NP;
    T *dinv = ptr; ptr += NP;
    T *spare = ptr; ptr += NP;
    T *red = ptr; ptr += 64;
    int *idx = reinterpret_cast<int *>(ptr);
    (void)spare;
    if (warp == 0) {
        int nf = 0;
        for (int base = 0; base < N; base += 32) {
            const int i = base + lane;
            const int v = i < N ? (int)p.state[b * N + i] : 2;
            const unsigned mf = __ballot_sync(kFullMask, v == 0);
            if (v == 0) idx[nf + __popc(mf & ((1u << lane) - 1u))] = i;
            nf += __popc(mf);
        }
    }
    for (int i = tid; i < N; i += kThreads) { wv[i] = p.w[b * N + i]; gv[i] = p.grad_w[b * N + i]; dv[i] = T(0); }
    const T *Lb;
    int ldl;
    if (kGlobal) {
        const T *Qg = p.saved + b * p.saved_stride;
        Lb = Qg + 1; ldl = p.LDF;
        for (int k = tid; k < nF; k += kThreads) dinv[k] = Qg[(long long)k * p.LDF + k];
    } else {
        const T *src = p.saved + b * p.saved_stride;
        ldl = nF | 1;
        for (int cc = warp; cc < nF; cc += kWarps) {
            const int off = cc * nF - cc * (cc - 1) / 2 - cc;
            for (int a = cc + lane; a < nF; a += 32) Lsm[cc * ldl + a] = src[off + a];
        }
        for (int k = tid; k < nF; k += kThreads) dinv[k] = src[nF * (nF + 1) / 2 + k];
        Lb = Lsm;
    }
    __syncthreads();
    if (warp == 0 && nF > 0) {
        T z1[NREG], z2[NREG];
#pragma unroll
        for (int m = 0; m < NREG; ++m) {
            const int k = lane + 32 * m;
            z1[m] = k < nF ? gv[idx[k]] : T(0);
            z2[m] = k < nF ? T(1) : T(0);
        }
        fwdsub2_warp<T, NREG>(Lb, ldl, nF, dinv, z1, z2);      // both right-hand sides through one pass over the factor
        T s12 = T(0), s22 = T(0);
#pragma unroll
        for (int m = 0; m < NREG; ++m) {
            const int k = lane + 32 * m;
            if (k < nF) { s12 = fma(z1[m] * z2[m], dinv[k], s12); s22 = fma(z2[m] * z2[m], dinv[k], s22); }
        }
        s12 = warp_sum(s12); s22 = warp_sum(s22);
        const T eta = s12 / s22;
#pragma unroll
        for (int m = 0; m < NREG; ++m) z1[m] = T(0.5) * (z1[m] - eta * z2[m]);
        backsub_warp<T, NREG>(Lb, ldl, nF, dinv, z1);
#pragma unroll
        for (int m = 0; m < NREG; ++m) {
            const int k = lane + 32 * m;
            if (k < nF) dv[idx[k]] = z1[m];
        }
    }
    __syncthreads();
    const T *Cb = p.C + b * (long long)N * N;
    const long long gflat0 = (p.b0 + b) * (long long)N * N;
    const T alpha = p.alpha[b];
    const unsigned uB = (unsigned)p.Bmod;
    T *pv = gv;   // p = A d overwrites grad_w now that d is known
    T ldot = T(0);
    for (int i = tid; i < N; i += kThreads) { p.d_rets[b * N + i] = dv[i]; ldot = fma(dv[i], wv[i], ldot); }
    const T dot = block_sum<T, kThreads>(ldot, red);
    if (tid == 0) p.d_alpha[b] = (alpha > T(0) ? T(1) : (alpha < T(0) ? T(-1) : T(0))) * (T(-2) * dot);
    if (nF > 0) {
        // p = A d, q = A w: one warp per row, coalesced reads of covmat_sqrt
        unsigned grow = (unsigned)((gflat0 + (long long)warp * N + lane) % p.Bmod);
        const unsigned growstep = (unsigned)(((long long)kWarps * N) % p.Bmod);
        const unsigned gstep = (unsigned)(32 % p.Bmod);
        for (int j = warp; j < N; j += kWarps) {
            T sp = T(0), sq = T(0);
            unsigned gi = grow;
            grow += growstep; if (grow >= uB) grow -= uB;
            for (int k = lane; k < N; k += 32) {
                const T a = Cb[(long long)j * N + k] * p.gamma[gi];
                sp = fma(a, dv[k], sp); sq = fma(a, wv[k], sq);
                gi += gstep; if (gi >= uB) gi -= uB;
            }
            sp = warp_sum(sp); sq = warp_sum(sq);
            if (lane == 0) { pv[j] = sp; qv[j] = sq; }
        }
    } else {
        for (int j = tid; j < N; j += kThreads) { pv[j] = T(0); qv[j] = T(0); }
    }
    __syncthreads();
    if (p.vecs) for (int i = tid; i < N; i += kThreads) { p.vecs[(b * 2) * N + i] = pv[i]; p.vecs[(b * 2 + 1) * N + i] = qv[i]; }
    T *dC = p.d_C + b * (long long)N * N;
    if ((N & 3) == 0) {
        unsigned grow = (unsigned)((gflat0 + (long long)warp * N + 4 * lane) % p.Bmod);
        const unsigned growstep = (unsigned)(((long long)kWarps * N) % p.Bmod);
        const unsigned gstep = (unsigned)(128 % p.Bmod);
        const bool gvec = (uB & 3u) == 0u;
        for (int j = warp; j < N; j += kWarps) {
            const T pj = T(-2) * pv[j], qj = T(-2) * qv[j];
            unsigned gi = grow;
            for (int k = 4 * lane; k < N; k += 128) {
                const Vec4<T> w4 = ld4(wv + k), d4 = ld4(dv + k);
                Vec4<T> o, g4;
                if (gvec) g4 = ld4(p.gamma + gi);
                else { unsigned g = gi;
#pragma unroll
                    for (int q = 0; q < 4; ++q) { g4.v[q] = p.gamma[g]; if (++g >= uB) g = 0; } }
#pragma unroll
                for (int q = 0; q < 4; ++q) o.v[q] = g4.v[q] * fma(pj, w4.v[q], qj * d4.v[q]);
                st4(dC + (long long)j * N + k, o);
                gi += gstep; if (gi >= uB) gi -= uB;
            }
            grow += growstep; if (grow >= uB) grow -= uB;
        }
    } else {
        const int total = N * N;
        int j = tid / N, k = tid - j * N;
        const int dj = kThreads / N, dkk = kThreads % N;
        unsigned gi = (unsigned)((gflat0 + tid) % p.Bmod);
        const unsigned gstep = (unsigned)(kThreads % p.Bmod);
        for (int e = tid; e < total; e += kThreads) {
            dC[e] = p.gamma[gi] * T(-2) * fma(pv[j], wv[k], qv[j] * dv[k]);
            k += dkk; j += dj;
            if (k >= N) { k -= N; ++j; }
            gi += gstep; if (gi >= uB) gi -= uB;
        }
    }
}

// d_gamma[m] += sum over the launch's elements with (flat % Bmod) == m of covmat_sqrt[flat] * dA[flat]
// (backward of the tiling).  `flat` counts from the first problem of the whole batch; this launch holds the
// `total` = B N^2 elements that start at flat = b0 N^2, so its local element f sits in bin (f + off) % Bmod with
// off = (b0 N^2) % Bmod.  grid: (ceil(bins / 128), nsplit); each thread owns kV neighbouring bins and a slice of
// q (local f = first + q Bmod); partial sums via atomics.  kV = 4 uses 128-bit loads and needs Bmod % 4 == 0
// and N % 4 == 0 (then off % 4 == 0 and the four elements f .. f+3 sit in one row of one problem).
template <typename T, int kV>
__global__ void __launch_bounds__(128) markowitz_gamma_kernel(const T *__restrict__ C, const T *__restrict__ w,
                                                              const T *__restrict__ d, const T *__restrict__ vecs,
                                                              long long total, long long Bmod, long long off, int N,
                                                              int q_per_split, T *d_gamma) {
    const long long m = kV * ((long long)blockIdx.x * 128 + threadIdx.x);
    if (m >= Bmod) return;
    const long long NN = (long long)N * N;
    long long first = m - off;
    if (first < 0) first += Bmod;
    const long long q0 = (long long)blockIdx.y * q_per_split;
    const long long flat0 = first + q0 * Bmod;
    if (flat0 >= total) return;
    long long b = flat0 / NN;
    int rem = (int)(flat0 - b * NN);
    int j = rem / N, k = rem - j * N;
    const long long Bq = Bmod / NN;
    const int Br = (int)(Bmod - Bq * NN), Brj = Br / N, Brk = Br - Brj * N;
    const int NNi = (int)NN;
    T acc[kV];
#pragma unroll
    for (int x = 0; x < kV; ++x) acc[x] = T(0);
    long long flat = flat0;
    // trip count known up front (no exit test on the address inside the loop): the loads of the unrolled iterations can be
    // issued together -- the kernel is pure memory latency (ncu: long_scoreboard 15 stalls per issue at 43% of the HBM peak)
    const long long left = (total - flat0 + Bmod - 1) / Bmod;
    const int nq = (int)(left < (long long)q_per_split ? left : (long long)q_per_split);
    auto advance = [&]() {
        flat += Bmod; b += Bq; rem += Br; j += Brj; k += Brk;
        if (k >= N) { k -= N; ++j; }
        if (rem >= NNi) { rem -= NNi; j -= N; ++b; }
    };
    int q = 0;
    if (kV == 4) {
        // four iterations at a time, every load issued before the first use (the index chain is integer work only)
        constexpr int kU = 4;
        for (; q + kU <= nq; q += kU) {
            long long bb[kU], ff[kU];
            int jj[kU], kk[kU];
#pragma unroll
            for (int t = 0; t < kU; ++t) { bb[t] = b; ff[t] = flat; jj[t] = j; kk[t] = k; advance(); }
            Vec4<T> c4[kU], w4[kU], d4[kU];
            T pj[kU], qj[kU];
#pragma unroll
            for (int t = 0; t < kU; ++t) {
                c4[t] = ld4(C + ff[t]); w4[t] = ld4(w + bb[t] * N + kk[t]); d4[t] = ld4(d + bb[t] * N + kk[t]);
                pj[t] = vecs[(bb[t] * 2) * N + jj[t]]; qj[t] = vecs[(bb[t] * 2 + 1) * N + jj[t]];
            }
#pragma unroll
            for (int t = 0; t < kU; ++t) {
                const T p2 = T(-2) * pj[t], q2 = T(-2) * qj[t];
#pragma unroll
                for (int x = 0; x < kV; ++x) acc[x] = fma(c4[t].v[x], fma(p2, w4[t].v[x], q2 * d4[t].v[x]), acc[x]);
            }
        }
    }
    for (; q < nq; ++q) {
        const T pj = T(-2) * vecs[(b * 2) * N + j], qj = T(-2) * vecs[(b * 2 + 1) * N + j];
        if (kV == 4) {
            const Vec4<T> c4 = ld4(C + flat), w4 = ld4(w + b * N + k), d4 = ld4(d + b * N + k);
#pragma unroll
            for (int x = 0; x < kV; ++x) acc[x] = fma(c4.v[x], fma(pj, w4.v[x], qj * d4.v[x]), acc[x]);
        } else {
            acc[0] = fma(C[flat], fma(pj, w[b * N + k], qj * d[b * N + k]), acc[0]);
        }
        advance();
    }
#pragma unroll
    for (int x = 0; x < kV; ++x) atomicAdd(d_gamma + m + x, acc[x]);
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
struct MkPlan {
    int threads, nreg, LD, AS, CH;
    bool global_q;
    size_t smem;
};

static int smem_limit() { return device_smem_limit(); }   // of the CURRENT device

static MkPlan plan_fwd(int N, int tsize) {
    MkPlan pl;
    pl.threads = N <= 128 ? 128 : (N <= 256 ? 256 : 512);
    pl.nreg = N <= 128 ? 4 : (N <= 256 ? 8 : (N <= 512 ? 16 : 32));
    pl.LD = (N + 3) | 1;
    pl.AS = round_up(N, 4);
    pl.CH = max(1, min(N, pl.threads * kPre / N));
    pl.global_q = false;
    pl.smem = mk_smem_bytes(N, pl.LD, pl.AS, pl.CH, tsize, false);
    if (pl.smem > (size_t)smem_limit()) {
        pl.global_q = true;
        pl.smem = mk_smem_bytes(N, pl.LD, pl.AS, pl.CH, tsize, true);
    }
    return pl;
}

static MkPlan plan_bwd(int N, int tsize) {
    MkPlan pl = plan_fwd(N, tsize);
    pl.LD = (N + 2) | 1;
    pl.global_q = false;
    pl.smem = mk_bwd_smem_bytes(N, pl.LD, pl.AS, pl.CH, tsize, false);
    if (pl.smem > (size_t)smem_limit()) {
        pl.global_q = true;
        pl.smem = mk_bwd_smem_bytes(N, pl.LD, pl.AS, pl.CH, tsize, true);
    }
    return pl;
}

constexpr int kIpmGrid = 296;

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

size_t markowitz_workspace_bytes(long long B, int N, int tsize);
size_t markowitz_saved_bytes(long long B, int N, int tsize);
static long long saved_stride_elems(int N, int tsize) {
    const MkPlan pf = plan_fwd(N, tsize);
    return pf.global_q ? (long long)N * pf.LD : (long long)round_up(N * (N + 1) / 2 + N, 4);
}
#ifdef DDW_MK_DEFINE_WORKSPACE
size_t markowitz_saved_bytes(long long B, int N, int tsize) {
    return align256(sizeof(int) * (size_t)B) + (size_t)B * (size_t)saved_stride_elems(N, tsize) * tsize;
}
size_t markowitz_workspace_bytes(long long B, int N, int tsize) {
    const MkPlan pf = plan_fwd(N, tsize), pb = plan_bwd(N, tsize);
    const int NP = round_up(N + 2, 32);
    size_t bytes = 0;
    bytes += align256(sizeof(int) * (size_t)(2 * B + 8));                  // fail/work counters + lists
    bytes += align256((size_t)kIpmGrid * 8 * NP * tsize);                  // interior-point vectors
    bytes += align256((size_t)B * 2 * N * tsize);                          // (A d, A w) for the gamma pass
    size_t q = 0;
    if (pf.global_q) q = (size_t)B * N * pf.LD * tsize;
    if (pb.global_q) q = max(q, (size_t)B * N * pb.LD * tsize);
    bytes += align256(q);
    if (fwd_qwarp_eligible(N, tsize)) bytes += align256(fwd_qwarp_scratch_bytes(B, N));     // Q scratch of the warp-per-portfolio forward
    return bytes;
}

#endif  // DDW_MK_DEFINE_WORKSPACE

template <typename T> struct WsLayout {
    int *fail_count, *fail_list, *work_count, *work_list, *bad_count, *bwd_counts;
    T *ipm, *vecs, *q;
    float *qs;        // Q scratch of markowitz_qwarp.cu (behind q)
};
template <typename T> static WsLayout<T> ws_layout(void *ws, long long B, int N) {
    const int NP = round_up(N + 2, 32);
    WsLayout<T> l;
    unsigned char *p = reinterpret_cast<unsigned char *>(ws);
    l.fail_count = reinterpret_cast<int *>(p);   // [0] fail_count, [1] work_count, [2] bad_count (cleared together)
    l.work_count = l.fail_count + 1;
    l.bad_count = l.fail_count + 2;
    l.bwd_counts = l.fail_count + 4;        // [4], [5]: the backward's two hand-over counters (the forward's header stays readable)
    l.fail_list = l.fail_count + 8;
    l.work_list = l.fail_list + B;
    p += align256(sizeof(int) * (size_t)(2 * B + 8));
    l.ipm = reinterpret_cast<T *>(p); p += align256((size_t)kIpmGrid * 8 * NP * sizeof(T));
    l.vecs = reinterpret_cast<T *>(p); p += align256((size_t)B * 2 * N * sizeof(T));
    l.q = reinterpret_cast<T *>(p);
    {
        const MkPlan pf = plan_fwd(N, sizeof(T)), pb = plan_bwd(N, sizeof(T));
        size_t q = 0;
        if (pf.global_q) q = (size_t)B * N * pf.LD * sizeof(T);
        if (pb.global_q) q = max(q, (size_t)B * N * pb.LD * sizeof(T));
        p += align256(q);
    }
    l.qs = reinterpret_cast<float *>(p);
    return l;
}

constexpr int kDenseListGrid = 4 * 148;

template <typename T, int kThreads, int NREG, bool kGlobal, int MAXT>
static int launch_fwd_variant(MkFwdParams<T> p, const MkPlan &pl, T *ipm_ws, bool sparse_first, cudaStream_t st, float *qscratch = nullptr) {
    auto kf = markowitz_fwd_kernel<T, kThreads, NREG, kGlobal, MAXT>;
    auto ki = markowitz_ipm_kernel<T, kThreads, NREG, kGlobal>;
    static DeviceOnce configured;     // per template instance and per device
    const int dev = current_device();
    if (configured.needs(dev)) {
        // opt in to the device maximum once per device: the same instance serves several n_assets
        if (cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit()) != cudaSuccess ||
            cudaFuncSetAttribute(ki, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit()) != cudaSuccess)
            return check_launch("cudaFuncSetAttribute(markowitz_fwd)");
        configured.mark(dev);
    }
    cudaMemsetAsync(p.fail_count, 0, 4 * sizeof(int), st);
    int rc;
    if (sparse_first) {
        // small-support portfolios are finished by the sparse kernel (its own translation unit); the rest land
        // on work_list
        bool done = false;
        rc = 0;
        if constexpr (sizeof(T) == 4) {
            if (qscratch) { rc = launch_fwd_qwarp(p, qscratch, st); done = true; }
        }
        if (!done) rc = launch_fwd_sparse<T>(p, st);
        if (rc) return rc;
        kf<<<(unsigned)min((long long)kDenseListGrid, p.B), kThreads, pl.smem, st>>>(p);
    } else {
        p.work_count = nullptr; p.work_list = nullptr;
        kf<<<(unsigned)p.B, kThreads, pl.smem, st>>>(p);
    }
    if ((rc = check_launch("markowitz_fwd_kernel"))) return rc;
    const unsigned grid = (unsigned)min((long long)kIpmGrid, p.B);
    ki<<<grid, kThreads, pl.smem, st>>>(p, ipm_ws, MAXT < 0 ? 1 : 0);
    return check_launch("markowitz_ipm_kernel");
}

template <typename T>
int markowitz_fwd_t(const void *rets, const void *C, const void *gamma, const void *alpha, double u, long long B,
                    int N, void *w, int8_t *state, int32_t *status, void *ws, size_t ws_bytes, cudaStream_t st,
                    long long Bmod, long long b0, void *saved) {
    if (ws_bytes < markowitz_workspace_bytes(B, N, sizeof(T)) || !ws) {
        set_error("ddw_markowitz_fwd: workspace too small (%zu < %zu)", ws_bytes, markowitz_workspace_bytes(B, N, sizeof(T)));
        return DDW_ERR_WORKSPACE;
    }
    const MkPlan pl = plan_fwd(N, sizeof(T));
    if (pl.smem > (size_t)smem_limit()) { set_error("ddw_markowitz_fwd: n_assets=%d needs %zu B of shared memory", N, pl.smem); return DDW_ERR_UNSUPPORTED; }
    WsLayout<T> l = ws_layout<T>(ws, B, N);
    MkFwdParams<T> p;
    p.rets = (const T *)rets; p.C = (const T *)C; p.gamma = (const T *)gamma; p.alpha = (const T *)alpha;
    p.u = (T)u; p.B = B; p.Bmod = Bmod; p.b0 = b0; p.N = N; p.LD = pl.LD; p.AS = pl.AS; p.CH = pl.CH;
    p.w = (T *)w; p.state = state; p.status = status; p.ws = l.q;
    p.fail_count = l.fail_count; p.fail_list = l.fail_list;
    p.work_count = l.work_count; p.work_list = l.work_list; p.bad_count = l.bad_count;
    p.saved_n = nullptr; p.saved = nullptr; p.saved_stride = 0;
    if (saved) {
        // keep the converged factors for ddw_markowitz_bwd_saved; with Q in the global workspace the solver works in place in
        // the caller's `saved` buffer instead of the workspace
        p.saved_n = reinterpret_cast<int *>(saved);
        p.saved = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(saved) + align256(sizeof(int) * (size_t)B));
        p.saved_stride = saved_stride_elems(N, sizeof(T));
        // -1 = no factor.  With the sparse-support / warp-per-portfolio kernels in front every portfolio's entry is written by the
        // kernel that finishes it (one graph node less per step); the dense kernel alone skips the portfolios the box decides
        if (!fwd_sparse_eligible(N, sizeof(T), smem_limit()) || pl.global_q || pl.threads != 128)
            cudaMemsetAsync(saved, 0xFF, sizeof(int) * (size_t)B, st);
        if (pl.global_q) { p.ws = p.saved; l.q = p.saved; }
    }
    // sparse-support kernel first whenever A (4 N^2 bytes) fits shared memory with room for >= 2 CTAs per SM
    const bool sparse_first = fwd_sparse_eligible(N, sizeof(T), smem_limit());
    if (!pl.global_q) {
        if (pl.threads == 128) {
            // Gram tiles kept in registers: 1, 2, 3 or 5 tiles of 4x4 per thread (n_assets <= 60 / 88 / 108 / 128)
            const int mt = gram_tiles_per_thread(N, 128);
            float *qs = sparse_first && fwd_qwarp_eligible(N, sizeof(T)) ? l.qs : nullptr;
            if (mt <= 1) return launch_fwd_variant<T, 128, 4, false, 1>(p, pl, l.ipm, sparse_first, st, qs);
            if (mt <= 2) return launch_fwd_variant<T, 128, 4, false, 2>(p, pl, l.ipm, sparse_first, st, qs);
            if (mt <= 3) return launch_fwd_variant<T, 128, 4, false, 3>(p, pl, l.ipm, sparse_first, st, qs);
            return launch_fwd_variant<T, 128, 4, false, 5>(p, pl, l.ipm, sparse_first, st, qs);
        }
        if (pl.threads == 256) return launch_fwd_variant<T, 256, 8, false, 0>(p, pl, l.ipm, false, st);
    }
    if (sizeof(T) == 4 && pl.threads >= 256) {
        // Q in the global workspace (n_assets beyond shared memory), float32: the Gram runs on tensor cores as a pre-pass
        int rc = markowitz_gram_tc((const float *)C, (const float *)gamma, (const float *)alpha, B, Bmod, b0, N, pl.LD, (float *)l.q, st);
        if (rc) return rc;
        if (pl.threads == 256) return launch_fwd_variant<T, 256, 8, true, -1>(p, pl, l.ipm, false, st);
        if (pl.nreg == 16) return launch_fwd_variant<T, 512, 16, true, -1>(p, pl, l.ipm, false, st);
        return launch_fwd_variant<T, 512, 32, true, -1>(p, pl, l.ipm, false, st);
    }
    if (pl.threads == 128) return launch_fwd_variant<T, 128, 4, true, 0>(p, pl, l.ipm, false, st);
    if (pl.threads == 256) return launch_fwd_variant<T, 256, 8, true, 0>(p, pl, l.ipm, false, st);
    if (pl.nreg == 16) return launch_fwd_variant<T, 512, 16, true, 0>(p, pl, l.ipm, false, st);
    return launch_fwd_variant<T, 512, 32, true, 0>(p, pl, l.ipm, false, st);
}

// ---- fused covariance -> NumericalMarkowitz forward ----------------------------------------
size_t markowitz_fused_smem_bytes(int N, int L, int tsize);
#ifdef DDW_MK_DEFINE_WORKSPACE
size_t markowitz_fused_smem_bytes(int N, int L, int tsize) {
    const MkPlan pl = plan_fwd(N, tsize);
    return (size_t)round_up((int)pl.smem, 16) + cov_tile_elems(L, N) * tsize;
}
#endif

template <typename T, int MAXT>
static int launch_fwd_fused(const MkFwdParams<T> &p, size_t smem, size_t smem_ipm, T *ipm_ws, cudaStream_t st) {
    auto kf = markowitz_fwd_kernel<T, 128, 4, false, MAXT, true>;
    auto ki = markowitz_ipm_kernel<T, 128, 4, false>;
    static DeviceOnce configured;
    const int dev = current_device();
    if (configured.needs(dev)) {
        if (cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit()) != cudaSuccess ||
            cudaFuncSetAttribute(ki, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit()) != cudaSuccess)
            return check_launch("cudaFuncSetAttribute(markowitz_fwd fused)");
        configured.mark(dev);
    }
    cudaMemsetAsync(p.fail_count, 0, 4 * sizeof(int), st);
    kf<<<(unsigned)p.B, 128, smem, st>>>(p);
    int rc = check_launch("markowitz_fwd_kernel (fused covariance)");
    if (rc || !p.C) return rc;
    // portfolios whose active set did not converge: their covariance tile was spilled to p.C, the fallback runs as usual
    ki<<<(unsigned)min((long long)kIpmGrid, p.B), 128, smem_ipm, st>>>(p, ipm_ws, 0);
    return check_launch("markowitz_ipm_kernel (fused covariance)");
}

// x (B, L, N) observations -> covariance (sqrt = False, `strategy`, coef) -> NumericalMarkowitz.  DDW_ERR_UNSUPPORTED when the
// tile does not fit next to the solver (n_assets > 128, long lookbacks): the caller composes the two layers instead.
template <typename T>
int markowitz_fused_fwd_t(const void *x, const void *coef, int strategy, int L, const void *rets, const void *gamma,
                          const void *alpha, double u, long long B, int N, void *w, int8_t *state, int32_t *status, void *ws,
                          size_t ws_bytes, cudaStream_t st, void *saved, void *cov_scratch) {
    const MkPlan pl = plan_fwd(N, sizeof(T));
    const size_t smem = markowitz_fused_smem_bytes(N, L, sizeof(T));
    if (pl.global_q || pl.threads != 128 || smem > (size_t)smem_limit() || L < 2) {
        set_error("ddw_covariance_markowitz_fwd: n_assets=%d, lookback=%d does not fit the fused kernel (%zu B of shared memory)", N, L, smem);
        return DDW_ERR_UNSUPPORTED;
    }
    if (ws_bytes < markowitz_workspace_bytes(B, N, sizeof(T)) || !ws) { set_error("ddw_covariance_markowitz_fwd: workspace too small"); return DDW_ERR_WORKSPACE; }
    WsLayout<T> l = ws_layout<T>(ws, B, N);
    MkFwdParams<T> p;
    memset(&p, 0, sizeof(p));
    p.rets = (const T *)rets; p.C = (const T *)cov_scratch; p.gamma = (const T *)gamma; p.alpha = (const T *)alpha;
    p.u = (T)u; p.B = B; p.Bmod = B; p.b0 = 0; p.N = N; p.LD = pl.LD; p.AS = pl.AS; p.CH = pl.CH;
    p.w = (T *)w; p.state = state; p.status = status; p.ws = l.q;
    p.fail_count = l.fail_count; p.fail_list = l.fail_list; p.bad_count = l.bad_count;
    p.X = (const T *)x; p.coef = (const T *)coef; p.L = L; p.strategy = strategy;
    if (saved) {
        p.saved_n = reinterpret_cast<int *>(saved);
        p.saved = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(saved) + align256(sizeof(int) * (size_t)B));
        p.saved_stride = saved_stride_elems(N, sizeof(T));
        cudaMemsetAsync(saved, 0xFF, sizeof(int) * (size_t)B, st);
    }
    const int mt = gram_tiles_per_thread(N, 128);
    if (mt <= 1) return launch_fwd_fused<T, 1>(p, smem, pl.smem, l.ipm, st);
    if (mt <= 2) return launch_fwd_fused<T, 2>(p, smem, pl.smem, l.ipm, st);
    if (mt <= 3) return launch_fwd_fused<T, 3>(p, smem, pl.smem, l.ipm, st);
    return launch_fwd_fused<T, 5>(p, smem, pl.smem, l.ipm, st);
}

template <typename T, int kThreads, int NREG, bool kGlobal>
static int launch_bwd_variant(const MkBwdParams<T> &p, const MkPlan &pl, cudaStream_t st) {
    auto kb = markowitz_bwd_kernel<T, kThreads, NREG, kGlobal>;
    static DeviceOnce configured;     // per template instance and per device
    const int dev = current_device();
    if (configured.needs(dev)) {
        if (cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit()) != cudaSuccess)
            return check_launch("cudaFuncSetAttribute(markowitz_bwd)");
        configured.mark(dev);
    }
    kb<<<(unsigned)p.B, kThreads, pl.smem, st>>>(p);
    return check_launch("markowitz_bwd_kernel");
}

template <typename T, int kThreads, int NREG, bool kGlobal>
static int launch_bwd_saved_variant(const MkBwdParams<T> &p, cudaStream_t st) {
    auto kb = markowitz_bwd_saved_kernel<T, kThreads, NREG, kGlobal>;
    static DeviceOnce configured;
    const int dev = current_device();
    if (configured.needs(dev)) {
        if (cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit()) != cudaSuccess)
            return check_launch("cudaFuncSetAttribute(markowitz_bwd_saved)");
        configured.mark(dev);
    }
    const size_t smem = mk_bwd_saved_smem_bytes(p.N, sizeof(T), kGlobal);
    if (smem > (size_t)smem_limit()) { set_error("ddw_markowitz_bwd_saved: n_assets=%d needs %zu B of shared memory", p.N, smem); return DDW_ERR_UNSUPPORTED; }
    kb<<<(unsigned)p.B, kThreads, smem, st>>>(p);
    return check_launch("markowitz_bwd_saved_kernel");
}

template <typename T>
int markowitz_bwd_t(const void *grad_w, const void *w, const int8_t *state, const void *C, const void *gamma,
                    const void *alpha, long long B, int N, void *d_rets, void *d_C, void *d_gamma, void *d_alpha,
                    void *ws, size_t ws_bytes, cudaStream_t st, long long Bmod, long long b0, int accumulate_gamma, const void *saved,
                    void *vecs_out) {
    if (ws_bytes < markowitz_workspace_bytes(B, N, sizeof(T)) || !ws) {
        set_error("ddw_markowitz_bwd: workspace too small (%zu < %zu)", ws_bytes, markowitz_workspace_bytes(B, N, sizeof(T)));
        return DDW_ERR_WORKSPACE;
    }
    const MkPlan pl = plan_bwd(N, sizeof(T));
    if (pl.smem > (size_t)smem_limit()) { set_error("ddw_markowitz_bwd: n_assets=%d needs %zu B of shared memory", N, pl.smem); return DDW_ERR_UNSUPPORTED; }
    WsLayout<T> l = ws_layout<T>(ws, B, N);
    MkBwdParams<T> p;
    p.grad_w = (const T *)grad_w; p.w = (const T *)w; p.C = (const T *)C; p.gamma = (const T *)gamma;
    p.alpha = (const T *)alpha; p.state = state; p.B = B; p.Bmod = Bmod; p.b0 = b0; p.N = N; p.LDW = pl.LD; p.AS = pl.AS; p.CH = pl.CH;
    // (A d, A w) per problem: scratch of the gamma pass, or -- vecs_out -- an output of its own (the factored gradient:
    // d_covmat_sqrt = -2 gamma_ * [(A d) w' + (A w) d'] is rank 2, 2 N numbers instead of N^2)
    T *vecs = vecs_out ? (T *)vecs_out : l.vecs;
    p.d_rets = (T *)d_rets; p.d_C = (T *)d_C; p.d_alpha = (T *)d_alpha; p.vecs = (d_gamma || vecs_out) ? vecs : nullptr; p.ws = l.q;
    int rc;
    // sparse-support kernel first (A staged by TMA, free set <= kCandMax); it lists the portfolios it cannot take
    p.work_count = nullptr; p.work_list = nullptr;
    p.saved_n = nullptr; p.saved = nullptr; p.saved_stride = 0; p.LDF = 0; p.work_count2 = nullptr; p.work_list2 = nullptr;
    cudaMemsetAsync(l.bwd_counts, 0, 2 * sizeof(int), st);          // both hand-over counters below, one graph node
    if (!pl.global_q && pl.threads == 128 && bwd_sparse_eligible(N, sizeof(T), smem_limit())) {
        p.work_count = l.bwd_counts + 1; p.work_list = l.work_list;
        if ((rc = launch_bwd_sparse<T>(p, st))) return rc;
    }
    const MkPlan pf = plan_fwd(N, sizeof(T));
    if (saved && mk_bwd_saved_smem_bytes(N, sizeof(T), pf.global_q) <= (size_t)smem_limit()) {
        // factors kept by the forward: substitutions only; what it cannot take (no factor saved) lands on the second list
        p.saved_n = reinterpret_cast<const int *>(saved);
        p.saved = reinterpret_cast<const T *>(reinterpret_cast<const unsigned char *>(saved) + align256(sizeof(int) * (size_t)B));
        p.saved_stride = saved_stride_elems(N, sizeof(T));
        p.LDF = pf.LD;
        p.work_count2 = l.bwd_counts; p.work_list2 = l.fail_list;
        if (!pf.global_q && pl.threads == 128) rc = launch_bwd_saved_variant<T, 128, 4, false>(p, st);
        else if (!pf.global_q) rc = launch_bwd_saved_variant<T, 256, 8, false>(p, st);
        else if (pl.threads == 256) rc = launch_bwd_saved_variant<T, 256, 8, true>(p, st);
        else if (pl.nreg == 16) rc = launch_bwd_saved_variant<T, 512, 16, true>(p, st);
        else rc = launch_bwd_saved_variant<T, 512, 32, true>(p, st);
        if (rc) return rc;
        p.work_count = p.work_count2; p.work_list = p.work_list2;      // the generic kernel walks what is left
        p.saved_n = nullptr;
    }
    if (!pl.global_q && pl.threads == 128) rc = launch_bwd_variant<T, 128, 4, false>(p, pl, st);
    else if (!pl.global_q && pl.threads == 256) rc = launch_bwd_variant<T, 256, 8, false>(p, pl, st);
    else if (pl.threads == 128) rc = launch_bwd_variant<T, 128, 4, true>(p, pl, st);
    else if (pl.threads == 256) rc = launch_bwd_variant<T, 256, 8, true>(p, pl, st);
    else if (pl.nreg == 16) rc = launch_bwd_variant<T, 512, 16, true>(p, pl, st);
    else rc = launch_bwd_variant<T, 512, 32, true>(p, pl, st);
    if (rc || !d_gamma) return rc;
    if (!accumulate_gamma) cudaMemsetAsync(d_gamma, 0, sizeof(T) * (size_t)Bmod, st);
    const long long NN = (long long)N * N, total = B * NN, off = (b0 * NN) % Bmod;
    const bool vec4 = (Bmod % 4 == 0) && (N % 4 == 0);
    const long long cols = vec4 ? Bmod / 4 : Bmod;   // threads along the bin dimension
    const long long qn = (total + Bmod - 1) / Bmod;  // elements per bin (upper bound)
    // enough slices to fill the machine, few enough that atomics stay negligible
    long long nsplit = max(1LL, min(min(qn, 65535LL), (148LL * 16 * 128 + cols - 1) / cols));
    const int q_per = (int)((qn + nsplit - 1) / nsplit);
    nsplit = (qn + q_per - 1) / q_per;
    dim3 grid((unsigned)((cols + 127) / 128), (unsigned)nsplit);
    if (vec4)
        markowitz_gamma_kernel<T, 4><<<grid, 128, 0, st>>>((const T *)C, (const T *)w, (const T *)d_rets, vecs, total,
                                                           Bmod, off, N, q_per, (T *)d_gamma);
    else
        markowitz_gamma_kernel<T, 1><<<grid, 128, 0, st>>>((const T *)C, (const T *)w, (const T *)d_rets, vecs, total,
                                                           Bmod, off, N, q_per, (T *)d_gamma);
    return check_launch("markowitz_gamma_kernel");
}

}  // namespace ddw
