// markowitz_sparse.cuh -- forward solver for portfolios with a small support (included by markowitz.cu)
//
// The dense kernel forms the whole Gram Q = A'A (N^3/2 flops) although an active-set iteration only
// ever needs (i) the block of Q on the assets that are or were free and (ii) gradients 2 A'(A w).
// Here A itself (gamma-tiled covmat_sqrt, 4 N^2 bytes = one contiguous block of HBM) is brought into
// shared memory by ONE bulk TMA copy per portfolio, and Q entries are computed only for assets when
// they first become free ("candidates").  Work per portfolio drops from ~N^3/2 to
// ~|cand|^2 N/2 + iterations * 2 N^2 flops.  Portfolios whose candidate set outgrows kCandMax are
// appended to a work list that the dense kernel then processes.
#pragma once

namespace ddw {

constexpr int kCandMax = 48;
constexpr int kCandLd = (kCandMax + 3) | 1;

template <typename T> struct SpSmem {
    unsigned long long *mbar;
    T *As, *Qc, *r, *wv, *tv, *g, *qd, *dinv, *red;
    int *idx, *idxU, *slot, *cand, *misc;
    int8_t *st;
};

static size_t sp_smem_bytes(int N, int tsize) {
    const int NP = round_up(N + 2, 32);
    const size_t elems = (size_t)round_up(N * N, 4) + round_up(kCandMax * kCandLd, 4) + 6 * (size_t)NP + 64;
    return 16 + elems * tsize + ((size_t)3 * NP + 64 + 8) * sizeof(int) + NP;
}

template <typename T> __device__ __forceinline__ SpSmem<T> sp_carve(unsigned char *raw, int N) {
    SpSmem<T> s;
    const int NP = round_up(N + 2, 32);
    s.mbar = reinterpret_cast<unsigned long long *>(raw);
    T *p = reinterpret_cast<T *>(raw + 16);
    s.As = p; p += round_up(N * N, 4);
    s.Qc = p; p += round_up(kCandMax * kCandLd, 4);
    s.r = p; p += NP;
    s.wv = p; p += NP;
    s.tv = p; p += NP;
    s.g = p; p += NP;
    s.qd = p; p += NP;
    s.dinv = p; p += NP;
    s.red = p; p += 64;
    int *q = reinterpret_cast<int *>(p);
    s.idx = q; q += NP;
    s.idxU = q; q += NP;
    s.slot = q; q += NP;
    s.cand = q; q += 64;
    s.misc = q; q += 8;
    s.st = reinterpret_cast<int8_t *>(q);
    return s;
}

template <typename T, int kThreads, int NREG>
__global__ void __launch_bounds__(kThreads) markowitz_fwd_sparse_kernel(MkFwdParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long long b = blockIdx.x;
    const int N = p.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int LDC = kCandLd;
    const T u = p.u;
    T *w_out = p.w + b * N;
    int8_t *st_out = p.state + b * N;
    if (trivial_cases<T, kThreads>(N, u, w_out, st_out, p.status + b, p.bad_count)) return;
    SpSmem<T> s = sp_carve<T>(smem_raw, N);
    T *As = s.As, *Qc = s.Qc, *Lb = s.Qc + 1;
    const T *Cb = p.C + b * (long long)N * N;
    const long long gflat0 = (p.b0 + b) * (long long)N * N;
    const unsigned bytes = (unsigned)(N * N * sizeof(T));
    const unsigned uB = (unsigned)p.Bmod;

    // ---- A -> shared memory: one bulk TMA copy (40 KB at N = 100), completion on an mbarrier
    const bool use_tma = (bytes & 15u) == 0u && (reinterpret_cast<unsigned long long>(Cb) & 15ull) == 0ull;
    if (use_tma) {
        if (tid == 0) { mbar_init(s.mbar, 1); fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) { mbar_expect_tx(s.mbar, bytes); tma_bulk_g2s(As, Cb, bytes, s.mbar); }
    }
    for (int i = tid; i < N; i += kThreads) { s.r[i] = p.rets[b * N + i]; s.slot[i] = -1; }
    if (tid == 0) { s.misc[3] = 0; }
    const T aabs = absT(p.alpha[b]);
    if (use_tma) mbar_wait(s.mbar, 0);
    else for (int e = tid; e < N * N; e += kThreads) As[e] = Cb[e];
    // ---- gamma tiling (allocate.py:268-270), in place; every thread touches the elements it will not
    // need before the next barrier
    if ((N & 3) == 0 && (uB & 3u) == 0u) {
        unsigned gi = (unsigned)((gflat0 + 4 * tid) % p.Bmod);
        const unsigned gstep = (unsigned)((4 * kThreads) % p.Bmod);
        for (int e = 4 * tid; e < N * N; e += 4 * kThreads) {
            Vec4<T> a = ld4(As + e);
            const Vec4<T> g4 = ld4(p.gamma + gi);
#pragma unroll
            for (int q = 0; q < 4; ++q) a.v[q] *= g4.v[q];
            st4(As + e, a);
            gi += gstep; if (gi >= uB) gi -= uB;
        }
    } else {
        unsigned gi = (unsigned)((gflat0 + tid) % p.Bmod);
        const unsigned gstep = (unsigned)(kThreads % p.Bmod);
        for (int e = tid; e < N * N; e += kThreads) {
            As[e] *= p.gamma[gi];
            gi += gstep; if (gi >= uB) gi -= uB;
        }
    }
    __syncthreads();
    // ---- diag(Q) = column norms + |alpha|
    for (int i = tid; i < N; i += kThreads) {
        T acc = aabs;
        for (int k = 0; k < N; ++k) { const T a = As[k * N + i]; acc = fma(a, a, acc); }
        s.qd[i] = acc;
    }
    __syncthreads();
    if (warp == 0) {
        T dmax;
        diag_init_warp<T, NREG>(s.qd, 1, s.r, s.st, N, u, &dmax);
        if (lane == 0) s.red[33] = dmax;
    }
    __syncthreads();
    const T dmin = T(4) * Num<T>::eps() * s.red[33] + Num<T>::tiny();
    const T tolw = T(8) * Num<T>::eps();
    int lifted = 0, converged = 0, bail = 0;
    for (int it = 0; it < kMaxPdas; ++it) {
        // (a) lists of free / capped assets; first-time free assets get a candidate slot
        if (warp == 0) {
            int nf = 0, nu = 0;
            const int nc0 = s.misc[3];
            int nc = nc0;
            for (int base = 0; base < N; base += 32) {
                const int i = base + lane;
                const int v = i < N ? (int)s.st[i] : 2;
                const unsigned lt = (1u << lane) - 1u;
                const unsigned mf = __ballot_sync(kFullMask, v == 0), mu = __ballot_sync(kFullMask, v == 1);
                if (v == 0) s.idx[nf + __popc(mf & lt)] = i;
                if (v == 1) s.idxU[nu + __popc(mu & lt)] = i;
                const bool need = v == 0 && s.slot[i] < 0;
                const unsigned mn = __ballot_sync(kFullMask, need);
                if (need) {
                    const int sl = nc + __popc(mn & lt);
                    if (sl < kCandMax) { s.slot[i] = sl; s.cand[sl] = i; }
                }
                nf += __popc(mf); nu += __popc(mu); nc += __popc(mn);
            }
            if (lane == 0) { s.misc[0] = nf; s.misc[1] = nu; s.misc[2] = nc0; s.misc[3] = nc; }
        }
        for (int i = tid; i < N; i += kThreads) s.wv[i] = s.st[i] > 0 ? u : T(0);
        __syncthreads();
        const int nF = s.misc[0], nU = s.misc[1], nc0 = s.misc[2], nc1 = s.misc[3];
        if (nc1 > kCandMax) { bail = 1; break; }
        // (b) Gram entries of the new candidates against all candidates: Q(s,c) = a_s . a_c
        if (nc1 > nc0) {
            const int base_cnt = nc0 * (nc0 + 1) / 2, tot = nc1 * (nc1 + 1) / 2 - base_cnt;
            for (int e = tid; e < tot; e += kThreads) {
                const int t = base_cnt + e;
                int sr = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
                while (sr * (sr + 1) / 2 > t) --sr;
                while ((sr + 1) * (sr + 2) / 2 <= t) ++sr;
                const int c = t - sr * (sr + 1) / 2;
                const int ia = s.cand[sr], ic = s.cand[c];
                T acc;
                if (sr == c) acc = s.qd[ia];
                else {
                    T a0 = T(0), a1 = T(0);
                    int k = 0;
                    for (; k + 1 < N; k += 2) {
                        a0 = fma(As[k * N + ia], As[k * N + ic], a0);
                        a1 = fma(As[(k + 1) * N + ia], As[(k + 1) * N + ic], a1);
                    }
                    if (k < N) a0 = fma(As[k * N + ia], As[k * N + ic], a0);
                    acc = a0 + a1;
                }
                Qc[sr * LDC + c] = acc;
            }
        }
        // (c) t_U = A (u 1_U)
        if (nU > 0) {
            for (int k = tid; k < N; k += kThreads) {
                T acc = T(0);
                for (int m = 0; m < nU; ++m) acc += As[k * N + s.idxU[m]];
                s.tv[k] = u * acc;
            }
        }
        __syncthreads();
        const T c = T(1) - u * (T)nU;
        int changed = 0;
        T nu_mult = T(0);
        if (nF > 0) {
            // (d) reduced KKT system: Q_FF from the candidate block, rhs r_F - 2 A_F' t_U and 1
            auto init = [&](int i, int j) -> T {
                const int cj = s.idx[j];
                if (i < nF) {
                    const int sa = s.slot[s.idx[i]], sc = s.slot[cj];
                    return sa >= sc ? Qc[sa * LDC + sc] : Qc[sc * LDC + sa];
                }
                if (i > nF) return T(1);
                T acc = s.r[cj];
                if (nU > 0) {
                    T d0 = T(0);
                    for (int k = 0; k < N; ++k) d0 = fma(As[k * N + cj], s.tv[k], d0);
                    acc -= T(2) * d0;
                }
                return acc;
            };
            if (small_system_fits<T, kThreads>(nF)) lifted |= factor_small<T, kThreads>(Lb, LDC, nF, s.dinv, dmin, init);
            else {
                for (int e = tid; e < nF * (nF + 2); e += kThreads) {
                    const int j = e / (nF + 2), i = e - j * (nF + 2);
                    if (i >= j) Lb[j * LDC + i] = init(i, j);
                }
                __syncthreads();
                lifted |= factor_ldl<T, kThreads>(Lb, LDC, nF, 2, s.dinv, dmin);
            }
            if (warp == 0) {
                T s12 = T(0), s22 = T(0);
                for (int k = lane; k < nF; k += 32) {
                    const T y1 = Lb[k * LDC + nF], y2 = Lb[k * LDC + nF + 1], di = s.dinv[k];
                    s12 = fma(y1 * y2, di, s12);
                    s22 = fma(y2 * y2, di, s22);
                }
                s12 = warp_sum(s12); s22 = warp_sum(s22);
                const T nu = (s12 - T(2) * c) / s22;
                T zt[NREG];
#pragma unroll
                for (int m = 0; m < NREG; ++m) {
                    const int k = lane + 32 * m;
                    zt[m] = k < nF ? T(0.5) * (Lb[k * LDC + nF] - nu * Lb[k * LDC + nF + 1]) : T(0);
                }
                backsub_warp<T, NREG>(Lb, LDC, nF, s.dinv, zt);
#pragma unroll
                for (int m = 0; m < NREG; ++m) {
                    const int k = lane + 32 * m;
                    if (k < nF) s.wv[s.idx[k]] = zt[m];
                }
                if (lane == 0) s.red[32] = nu;
            }
            __syncthreads();
            nu_mult = s.red[32];
        }
        // (e) t = A w on the support, then the gradient 2 (A' t + |alpha| w) - r + nu
        for (int k = tid; k < N; k += kThreads) {
            T acc = nU > 0 ? s.tv[k] : T(0);
            for (int m = 0; m < nF; ++m) { const int j = s.idx[m]; acc = fma(As[k * N + j], s.wv[j], acc); }
            s.tv[k] = acc;
        }
        __syncthreads();
        for (int i = tid; i < N; i += kThreads) {
            T a0 = T(0), a1 = T(0);
            int k = 0;
            for (; k + 3 < N; k += 4) {
                const Vec4<T> t4 = ld4(s.tv + k);
                a0 = fma(As[k * N + i], t4.v[0], a0);
                a1 = fma(As[(k + 1) * N + i], t4.v[1], a1);
                a0 = fma(As[(k + 2) * N + i], t4.v[2], a0);
                a1 = fma(As[(k + 3) * N + i], t4.v[3], a1);
            }
            for (; k < N; ++k) a0 = fma(As[k * N + i], s.tv[k], a0);
            const T h = T(2) * (a0 + a1 + aabs * s.wv[i]) - s.r[i];
            if (nF > 0) {
                const T gi = h + nu_mult;
                const int8_t so = s.st[i];
                int8_t sn = so;
                if (so == 0) {
                    const T wi = s.wv[i];
                    if (wi < -tolw) sn = -1;
                    else if (wi > u + tolw) sn = 1;
                } else if (so < 0) { if (gi < T(0)) sn = 0; }
                else { if (gi > T(0)) sn = 0; }
                if (sn != so) { s.st[i] = sn; changed = 1; }
            } else s.g[i] = h;
        }
        if (nF == 0) {
            // no free asset: optimal iff max_L(-h) <= min_U(-h) and u |U| = 1 (same rule as the dense kernel)
            __syncthreads();
            if (warp == 0) {
                T hl = Num<T>::inf(), hu = Num<T>::inf();
                int il = 0x7fffffff, iu = 0x7fffffff;
                for (int i = lane; i < N; i += 32) {
                    const T h = s.g[i];
                    if (s.st[i] < 0) { if (h < hl) { hl = h; il = i; } }
                    else if (s.st[i] > 0) { if (-h < hu) { hu = -h; iu = i; } }
                }
                warp_argmin(hl, il); warp_argmin(hu, iu);
                const T ctol = T(4) * Num<T>::eps() * (T)N;
                if (lane == 0) {
                    int ch = 0;
                    if (c > ctol && il < N) { s.st[il] = 0; ch = 1; }
                    else if (c < -ctol && iu < N) { s.st[iu] = 0; ch = 1; }
                    else if (il < N && iu < N && -hl > hu) { s.st[il] = 0; s.st[iu] = 0; ch = 1; }
                    s.misc[4] = ch;
                }
            }
            __syncthreads();
            changed = s.misc[4];
        }
        if (!__syncthreads_or(changed)) { converged = 1; break; }
    }
    if (bail) {
        // support too large for this kernel: hand the portfolio to the dense kernel
        if (tid == 0) { const int sl = atomicAdd(p.work_count, 1); p.work_list[sl] = (int)b; }
        return;
    }
    for (int i = tid; i < N; i += kThreads) {
        int8_t st = s.st[i];
        const T wi = st == 0 ? clampT(s.wv[i], T(0), u) : (st > 0 ? u : T(0));
        if (st == 0) st = wi <= T(0) ? -1 : (wi >= u ? 1 : 0);   // tie rule, see write_result
        w_out[i] = wi;
        st_out[i] = st;
    }
    if (tid == 0) {
        int st = lifted ? DDW_STATUS_REGULARIZED : 0;
        if (!converged) {
            st |= DDW_STATUS_FALLBACK | DDW_STATUS_INEXACT;
            const int sl = atomicAdd(p.fail_count, 1);
            p.fail_list[sl] = (int)b;
        }
        p.status[b] = st;
    }
}

}  // namespace ddw
