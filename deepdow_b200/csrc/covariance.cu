// covariance.cu -- batched sample covariance + shrinkage + symmetric PSD square root, fwd and bwd.
//
// Replaces the per-sample python loop of CovarianceMatrix.forward (reference
// deepdow/layers/misc.py:107-119): compute_covariance (:143-166) and compute_sqrt (:183-201, an SVD
// per sample).  One CTA per sample; the N x N matrix never leaves shared memory between the
// centred Gram, the shrinkage epilogue, the eigen-decomposition and the reconstruction.
//
// Square root: parallel-order cyclic Jacobi on the symmetric matrix (round-robin pairing, N/2
// disjoint rotations per round), eigenvectors accumulated as rows of VT.  The reference takes
// v * sqrt(s) * v' from an SVD, i.e. sqrt|lambda| with singular values below s.max * N * eps
// dropped (:185-199); so do we.
//
// Backward of the square root solves the Sylvester equation R dS + dS R = sym(G) in the
// eigenbasis, dS = V [(V' G V) / (r_i + r_j)] V', which stays finite for repeated eigenvalues
// (the reference's SVD autograd returns NaN there).
#include "common.cuh"
#include "linalg.cuh"
#include "cov_tile.cuh"

namespace ddw {

constexpr int kMaxSweeps = 30;

template <typename T> struct CovParams {
    const T *x, *coef, *grad_out, *eigvec_in, *eigval_in;
    int strategy, do_sqrt;
    long long B;
    int L, N, LDS, LDV, AS, CH;
    T *out, *eigvec, *eigval, *d_x, *d_coef;
    int32_t *sweeps;
    T *ws;
};

// shrinkage target weight bookkeeping (misc.py:149-166), applied to the lower triangle in place:
// S <- c S + (1 - c) T(S)
template <typename T, int kThreads>
__device__ __forceinline__ void shrink_lower(T *S, int LDS, int N, int strategy, T c, T fact, T *red) {
    const int tid = threadIdx.x;
    T tr = T(0);
    if (strategy == DDW_SHRINK_SCALED_IDENTITY) {
        T loc = T(0);
        for (int i = tid; i < N; i += kThreads) loc += S[i * LDS + i] * fact;
        tr = block_sum<T, kThreads>(loc, red) / (T)N;
    }
    const int tot = N * (N + 1) / 2;
    for (int e = tid; e < tot; e += kThreads) {
        int i = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
        while (i * (i + 1) / 2 > e) --i;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        const int j = e - i * (i + 1) / 2;
        const T s = S[i * LDS + j] * fact;
        T v = s;
        if (strategy == DDW_SHRINK_DIAGONAL) v = i == j ? c * s + (T(1) - c) * s : c * s;
        else if (strategy == DDW_SHRINK_IDENTITY) v = c * s + (i == j ? (T(1) - c) : T(0));
        else if (strategy == DDW_SHRINK_SCALED_IDENTITY) v = c * s + (i == j ? (T(1) - c) * tr : T(0));
        S[i * LDS + j] = v;
    }
    __syncthreads();
}

// write the full symmetric matrix held as a lower triangle
template <typename T, int kThreads>
__device__ __forceinline__ void store_sym(const T *S, int LDS, int N, T *out) {
    int i = threadIdx.x / N, j = threadIdx.x - i * N;
    const int di = kThreads / N, dj = kThreads % N;
    for (int e = threadIdx.x; e < N * N; e += kThreads) {
        out[e] = i >= j ? S[i * LDS + j] : S[j * LDS + i];
        j += dj; i += di;
        if (j >= N) { j -= N; ++i; }
    }
}

template <typename T, int kThreads, bool kGlobal>
__global__ void __launch_bounds__(kThreads) covariance_fwd_kernel(CovParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long long b = blockIdx.x;
    const int N = p.N, L = p.L, LDS = p.LDS, LDV = p.LDV, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;
    const int NP = round_up(N + 2, 32);
    T *ptr = reinterpret_cast<T *>(smem_raw);
    T *S, *VT = nullptr;
    if (kGlobal) {
        S = p.ws + b * ((long long)round_up(N * LDS, 4) + (long long)round_up(N * LDV, 4));
        VT = S + round_up(N * LDS, 4);
    } else {
        S = ptr; ptr += round_up(N * LDS, 4);
        if (p.do_sqrt) { VT = ptr; ptr += round_up(N * LDV, 4); }
    }
    T *As = ptr; ptr += 2 * p.CH * p.AS;
    T *mean = ptr; ptr += NP;
    T *rs = ptr; ptr += NP;
    T *cs = ptr; ptr += NP;
    T *sn = ptr; ptr += NP;
    T *red = ptr; ptr += 64;
    int *pp = reinterpret_cast<int *>(ptr);
    int *pq = pp + NP;

    const bool direct = L == 0;   // ddw_psd_sqrt_fwd: x already is the (B, N, N) symmetric matrix
    const T *xb = p.x + b * (long long)(direct ? N : L) * N;
    if (direct) {
        for (int i = warp; i < N; i += kWarps)
            for (int j = lane; j <= i; j += 32) S[i * LDS + j] = T(0.5) * (xb[(long long)i * N + j] + xb[(long long)j * N + i]);
        __syncthreads();
    } else {
        // 1. column means
        for (int i = tid; i < N; i += kThreads) {
            T acc = T(0);
            for (int l = 0; l < L; ++l) acc += xb[(long long)l * N + i];
            mean[i] = acc / (T)L;
        }
        __syncthreads();
        // 2. centred Gram (lower triangle, unscaled)
        gram_lower<T, kThreads, false>(xb, nullptr, 0, 1, L, N, mean, nullptr, N, As, p.AS, p.CH, S, LDS, T(0));
        // 3. 1/(L-1) and shrinkage
        const T fact = T(1) / (T)(L - 1);
        const T c = p.strategy == DDW_SHRINK_NONE ? T(1) : p.coef[b];
        shrink_lower<T, kThreads>(S, LDS, N, p.strategy, c, fact, red);
    }
    T *outb = p.out + b * (long long)N * N;
    if (!p.do_sqrt) {
        store_sym<T, kThreads>(S, LDS, N, outb);
        return;
    }
    // 4. Jacobi eigen-decomposition of the full symmetric matrix
    for (int e = tid; e < N * LDV; e += kThreads) VT[e] = T(0);
    for (int i = warp; i < N; i += kWarps)
        for (int j = i + 1 + lane; j < N; j += 32) S[i * LDS + j] = S[j * LDS + i];
    __syncthreads();
    for (int i = tid; i < N; i += kThreads) VT[i * LDV + i] = T(1);
    const int ne = N + (N & 1), npairs = ne / 2, nm1 = ne - 1;
    const T tol = Num<T>::eps();
    int sweep = 0;
    __syncthreads();
    for (; sweep < kMaxSweeps && N > 1; ++sweep) {
        int rotated = 0;
        for (int r = 0; r < nm1; ++r) {
            for (int k = tid; k < npairs; k += kThreads) {
                int a, bb;
                if (k == 0) { a = r; bb = nm1; }
                else { a = (r + k) % nm1; bb = (r - k + nm1) % nm1; }
                const int pi = min(a, bb), qi = max(a, bb);
                T cc = T(1), ss = T(0);
                if (qi < N) {
                    const T spq = S[pi * LDS + qi], spp = S[pi * LDS + pi], sqq = S[qi * LDS + qi];
                    const T aq = absT(spq);
                    if (aq > tol * sqrtT(absT(spp * sqq)) && aq > Num<T>::tiny()) {
                        const T tau = (sqq - spp) / (T(2) * spq);
                        const T t = (tau >= T(0) ? T(1) : T(-1)) / (absT(tau) + sqrtT(T(1) + tau * tau));
                        cc = T(1) / sqrtT(T(1) + t * t);   // correctly rounded: rsqrt's bias would shrink V over ~1e3 rotations
                        ss = t * cc;
                        rotated = 1;
                    }
                }
                // Rutishauser's form a' = a - s (b + tau a), b' = b + s (a - tau b), tau = s / (1 + c):
                // unlike c*a - s*b it does not lose the 1 - c term when 1 + t^2 rounds to 1, which
                // would otherwise stretch the eigenvectors by ~t^2/2 on every small rotation
                pp[k] = pi; pq[k] = qi; cs[k] = ss / (T(1) + cc); sn[k] = ss;
            }
            __syncthreads();
            // columns p,q of S and rows p,q of VT
            if (kGlobal) {
                // matrix in the global workspace: walk S row by row -- a warp owns row i and its lanes take the
                // (disjoint) pairs, so every access stays inside one contiguous row.  The column-wise walk below
                // touches 2 x 32 cache lines per instruction there (measured at N=500, 1024 samples: 15.9 s -> 9.2 s;
                // what remains is the ~24 GB per sample that 4000 rotation rounds stream through L2 -- a blocked
                // Jacobi is the next step for n_assets beyond shared memory)
                for (int i = warp; i < N; i += kWarps) {
                    T *row = S + i * LDS;
                    for (int k = lane; k < npairs; k += 32) {
                        const T ss = sn[k];
                        if (ss == T(0)) continue;
                        const T tt = cs[k];
                        const int pi = pp[k], qi = pq[k];
                        const T a = row[pi], c2 = row[qi];
                        row[pi] = fma(-ss, fma(a, tt, c2), a);
                        row[qi] = fma(ss, fma(-c2, tt, a), c2);
                    }
                }
            }
            for (int k = warp; k < npairs; k += kWarps) {
                const T ss = sn[k];
                if (ss == T(0)) continue;
                const T tt = cs[k];
                const int pi = pp[k], qi = pq[k];
                for (int i = lane; i < N; i += 32) {
                    if (!kGlobal) {
                        const T a = S[i * LDS + pi], c2 = S[i * LDS + qi];
                        S[i * LDS + pi] = fma(-ss, fma(a, tt, c2), a);
                        S[i * LDS + qi] = fma(ss, fma(-c2, tt, a), c2);
                    }
                    const T va = VT[pi * LDV + i], vb = VT[qi * LDV + i];
                    VT[pi * LDV + i] = fma(-ss, fma(va, tt, vb), va);
                    VT[qi * LDV + i] = fma(ss, fma(-vb, tt, va), vb);
                }
            }
            __syncthreads();
            // rows p,q of S
            for (int k = warp; k < npairs; k += kWarps) {
                const T ss = sn[k];
                if (ss == T(0)) continue;
                const T tt = cs[k];
                const int pi = pp[k], qi = pq[k];
                for (int j = lane; j < N; j += 32) {
                    const T a = S[pi * LDS + j], c2 = S[qi * LDS + j];
                    S[pi * LDS + j] = fma(-ss, fma(a, tt, c2), a);
                    S[qi * LDS + j] = fma(ss, fma(-c2, tt, a), c2);
                }
            }
            __syncthreads();
        }
        if (!__syncthreads_or(rotated)) { ++sweep; break; }
    }
    if (p.sweeps && tid == 0) p.sweeps[b] = sweep;
    // 5. sqrt|lambda| with the reference's rank threshold, then R = sum_k r_k v_k v_k'
    T smax = T(0);
    for (int i = tid; i < N; i += kThreads) smax = max(smax, absT(S[i * LDS + i]));
    smax = warp_max(smax);
    if (lane == 0) red[warp] = smax;
    __syncthreads();
    smax = T(0);
    for (int i = 0; i < kWarps; ++i) smax = max(smax, red[i]);
    const T thr = smax * (T)N * Num<T>::eps();
    for (int i = tid; i < N; i += kThreads) {
        const T sv = absT(S[i * LDS + i]);
        rs[i] = sv > thr ? sqrtT(sv) : T(0);
    }
    __syncthreads();
    {
        const int NT = (N + 3) >> 2, ntiles = NT * (NT + 1) / 2;
        for (int t = tid; t < ntiles; t += kThreads) {
            int ti = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
            while (ti * (ti + 1) / 2 > t) --ti;
            while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
            const int tj = t - ti * (ti + 1) / 2;
            T acc[4][4];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = T(0);
            for (int k = 0; k < N; ++k) {
                const Vec4<T> va = ld4(VT + k * LDV + 4 * ti), vb = ld4(VT + k * LDV + 4 * tj);
                const T rk = rs[k];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const T ax = va.v[x] * rk;
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = fma(ax, vb.v[y], acc[x][y]);
                }
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const int i = 4 * ti + x, j = 4 * tj + y;
                    if (i < N && j <= i) S[i * LDS + j] = acc[x][y];
                }
        }
    }
    __syncthreads();
    store_sym<T, kThreads>(S, LDS, N, outb);
    if (p.eigvec) {
        T *ev = p.eigvec + b * (long long)N * N;
        int i = tid / N, j = tid - i * N;
        const int di = kThreads / N, dj = kThreads % N;
        for (int e = tid; e < N * N; e += kThreads) {
            ev[e] = VT[i * LDV + j];
            j += dj; i += di;
            if (j >= N) { j -= N; ++i; }
        }
    }
    if (p.eigval) for (int i = tid; i < N; i += kThreads) p.eigval[b * N + i] = rs[i];
}

// C(i,j) = sum_k A[i*sai + k*sak] * B[k*sbk + j*sbj], i,j,k < n; each warp owns 4 rows, lanes own
// columns j = lane + 32 y.  Conflict-free for unit or odd strides.  Ends with a barrier.
template <typename T, int kThreads>
__device__ __forceinline__ void mm_smem(const T *A, int sai, int sak, const T *Bm, int sbk, int sbj, T *Cm, int ldc,
                                        int n) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kThreads / 32;
    const int nib = (n + 3) >> 2, njb = (n + 127) >> 7;
    for (int t = warp; t < nib * njb; t += kWarps) {
        const int ib = t / njb, jb = t - ib * njb;
        T acc[4][4];
        int jj[4];
        int ii[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            ii[x] = min(4 * ib + x, n - 1);
            jj[x] = min(jb * 128 + lane + 32 * x, n - 1);
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = T(0);
        }
        for (int k = 0; k < n; ++k) {
            T a[4], bv[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) { a[x] = A[ii[x] * sai + k * sak]; bv[x] = Bm[k * sbk + jj[x] * sbj]; }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], bv[y], acc[x][y]);
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const int i = 4 * ib + x, j = jb * 128 + lane + 32 * y;
                if (i < n && j < n) Cm[i * ldc + j] = acc[x][y];
            }
    }
    __syncthreads();
}

template <typename T, int kThreads, bool kGlobal>
__global__ void __launch_bounds__(kThreads) covariance_bwd_kernel(CovParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long long b = blockIdx.x;
    const int N = p.N, L = p.L, LD = p.LDS, tid = threadIdx.x;
    const int NP = round_up(N + 2, 32);
    T *ptr = reinterpret_cast<T *>(smem_raw);
    T *Gm, *S0, *VT = nullptr;
    const size_t msz = (size_t)round_up(N * LD, 4);
    if (kGlobal) {
        Gm = p.ws + b * (long long)(3 * msz);
        S0 = Gm + msz; VT = S0 + msz;
    } else {
        Gm = ptr; ptr += msz;
        S0 = ptr; ptr += msz;
        if (p.do_sqrt) { VT = ptr; ptr += msz; }
    }
    T *As = ptr; ptr += 2 * p.CH * p.AS;
    T *mean = ptr; ptr += NP;
    T *rs = ptr; ptr += NP;
    T *red = ptr; ptr += 64;

    const bool direct = L == 0;   // ddw_psd_sqrt_bwd: gradient w.r.t. the matrix itself
    const T *xb = p.x + b * (long long)L * N;
    const T *gb = p.grad_out + b * (long long)N * N;
    // symmetrised upstream gradient
    {
        int i = tid / N, j = tid - i * N;
        const int di = kThreads / N, dj = kThreads % N;
        for (int e = tid; e < N * N; e += kThreads) {
            Gm[i * LD + j] = T(0.5) * (gb[e] + gb[(long long)j * N + i]);
            if (p.do_sqrt) VT[i * LD + j] = p.eigvec_in[b * (long long)N * N + e];
            j += dj; i += di;
            if (j >= N) { j -= N; ++i; }
        }
    }
    for (int i = tid; i < N; i += kThreads) {
        if (!direct) {
            T acc = T(0);
            for (int l = 0; l < L; ++l) acc += xb[(long long)l * N + i];
            mean[i] = acc / (T)L;
        }
        if (p.do_sqrt) rs[i] = p.eigval_in[b * N + i];
    }
    __syncthreads();
    if (p.do_sqrt) {
        // G <- V [(V' G V) / (r_i + r_j)] V'   with V' = VT
        mm_smem<T, kThreads>(VT, LD, 1, Gm, LD, 1, S0, LD, N);      // S0 = VT G
        mm_smem<T, kThreads>(S0, LD, 1, VT, 1, LD, Gm, LD, N);      // Gm = S0 VT' = VT G V
        for (int e = tid; e < N * N; e += kThreads) {
            const int i = e / N, j = e - i * N;
            const T den = rs[i] + rs[j];
            Gm[i * LD + j] = den > T(0) ? Gm[i * LD + j] / den : T(0);
        }
        __syncthreads();
        mm_smem<T, kThreads>(Gm, LD, 1, VT, LD, 1, S0, LD, N);      // S0 = M VT
        mm_smem<T, kThreads>(VT, 1, LD, S0, LD, 1, Gm, LD, N);      // Gm = VT' S0 = V M V'
    }
    if (direct) {
        T *dm = p.d_x + b * (long long)N * N;
        int i = tid / N, j = tid - i * N;
        const int di = kThreads / N, dj = kThreads % N;
        for (int e = tid; e < N * N; e += kThreads) {
            dm[e] = Gm[i * LD + j];
            j += dj; i += di;
            if (j >= N) { j -= N; ++i; }
        }
        return;
    }
    // sample covariance (lower, unscaled) for the coefficient gradient
    const T fact = T(1) / (T)(L - 1);
    const T c = p.strategy == DDW_SHRINK_NONE ? T(1) : p.coef[b];
    T trG = T(0), trS = T(0);
    if (p.strategy != DDW_SHRINK_NONE) {
        gram_lower<T, kThreads, false>(xb, nullptr, 0, 1, L, N, mean, nullptr, N, As, p.AS, p.CH, S0, LD, T(0));
        T lg = T(0), ls = T(0), lgs = T(0), lgd = T(0);
        for (int e = tid; e < N * N; e += kThreads) {
            const int i = e / N, j = e - i * N;
            const T s = (i >= j ? S0[i * LD + j] : S0[j * LD + i]) * fact;
            const T g = Gm[i * LD + j];
            lgs = fma(g, s, lgs);
            if (i == j) { lg += g; ls += s; lgd = fma(g, s, lgd); }
        }
        trG = block_sum<T, kThreads>(lg, red);
        trS = block_sum<T, kThreads>(ls, red);
        const T gs = block_sum<T, kThreads>(lgs, red);
        const T gd = block_sum<T, kThreads>(lgd, red);
        T dc = T(0);
        if (p.strategy == DDW_SHRINK_IDENTITY) dc = gs - trG;
        else if (p.strategy == DDW_SHRINK_SCALED_IDENTITY) dc = gs - trS / (T)N * trG;
        else dc = gs - gd;
        if (p.d_coef && tid == 0) p.d_coef[b] = dc;
    }
    // D = 2 fact dS0
    for (int e = tid; e < N * N; e += kThreads) {
        const int i = e / N, j = e - i * N;
        T g = Gm[i * LD + j];
        if (p.strategy == DDW_SHRINK_IDENTITY) g = c * g;
        else if (p.strategy == DDW_SHRINK_SCALED_IDENTITY) g = c * g + (i == j ? (T(1) - c) * trG / (T)N : T(0));
        else if (p.strategy == DDW_SHRINK_DIAGONAL) g = i == j ? c * g + (T(1) - c) * g : c * g;
        Gm[i * LD + j] = T(2) * fact * g;
    }
    __syncthreads();
    // d_x[l, j] = sum_i xc[l, i] D[i, j]; four observation rows at a time, staged transposed
    T *xs = As;  // [N][4]
    T *dxb = p.d_x + b * (long long)L * N;
    for (int l0 = 0; l0 < L; l0 += 4) {
        for (int e = tid; e < 4 * N; e += kThreads) {
            const int r = e / N, i = e - r * N;
            xs[i * 4 + r] = (l0 + r < L) ? xb[(long long)(l0 + r) * N + i] - mean[i] : T(0);
        }
        __syncthreads();
        for (int j = tid; j < N; j += kThreads) {
            T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
            for (int i = 0; i < N; ++i) {
                const T d = Gm[i * LD + j];
                const Vec4<T> xv = ld4(xs + 4 * i);
                a0 = fma(xv.v[0], d, a0); a1 = fma(xv.v[1], d, a1);
                a2 = fma(xv.v[2], d, a2); a3 = fma(xv.v[3], d, a3);
            }
            dxb[(long long)l0 * N + j] = a0;
            if (l0 + 1 < L) dxb[(long long)(l0 + 1) * N + j] = a1;
            if (l0 + 2 < L) dxb[(long long)(l0 + 2) * N + j] = a2;
            if (l0 + 3 < L) dxb[(long long)(l0 + 3) * N + j] = a3;
        }
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------------
// CovarianceMatrix(sqrt=False) for samples that fit shared memory (BachelierNet's covariance, nn.py:170-171): one CTA per
// sample, the (L, N) lookback tile staged by ONE bulk TMA copy, centred into a padded copy, the Gram formed as 4x4 register
// tiles over the lower triangle with 16-byte shared loads, scale + shrinkage applied on the tile, the symmetric result
// assembled in shared memory and streamed out flat with 16-byte stores.  HBM traffic is the algorithmic
// (L N + N^2 + 1) elements per sample.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kPlainThreads) covariance_plain_kernel(CovParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long mbar;
    const long long b = blockIdx.x;
    const int N = p.N, L = p.L, XS = p.AS, tid = threadIdx.x;
    T *X = reinterpret_cast<T *>(smem_raw);                 // (L, N) dense, as in HBM
    T *Xc = X + round_up(L * N, 4);                          // (L, XS) centred, columns N..XS-1 zero
    T *S = Xc + L * XS;                                      // (N, N) dense result
    T *part = S + round_up(N * N, 4);                        // (kPlainParts, N) partial column sums
    T *red = part + kPlainParts * round_up(N, 4);
    const T *xb = p.x + b * (long long)L * N;
    const unsigned bytes = (unsigned)(L * N * sizeof(T));
    const bool use_tma = (bytes & 15u) == 0u && (reinterpret_cast<unsigned long long>(xb) & 15ull) == 0ull;
    if (use_tma) {
        if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) { mbar_expect_tx(&mbar, bytes); tma_bulk_g2s(X, xb, bytes, &mbar); }
        mbar_wait(&mbar, 0);
    } else {
        for (int e = tid; e < L * N; e += kPlainThreads) X[e] = xb[e];
        __syncthreads();
    }
    cov_plain_tile<T>(X, Xc, S, part, red, L, N, XS, p.strategy, p.strategy == DDW_SHRINK_NONE ? T(1) : p.coef[b]);
    T *outb = p.out + b * (long long)N * N;
    const int nn = N * N;
    if ((nn & 3) == 0 && (reinterpret_cast<unsigned long long>(outb) & (4 * sizeof(T) - 1)) == 0ull) {
        for (int e = 4 * tid; e < nn; e += 4 * kPlainThreads) st4(outb + e, ld4(S + e));
    } else {
        for (int e = tid; e < nn; e += kPlainThreads) outb[e] = S[e];
    }
}

static size_t cov_plain_smem(int L, int N, int tsize) {
    const int XS = round_up(N, 4);
    return ((size_t)round_up(L * N, 4) + (size_t)L * XS + (size_t)round_up(N * N, 4) + (size_t)kPlainParts * XS + 64) * tsize;
}
// at least 3 samples resident per SM, else the generic kernel (which streams the lookback rows) is the better fit
static bool cov_plain_fits(int L, int N, int tsize) { return L >= 2 && cov_plain_smem(L, N, tsize) <= 75 * 1024; }

// ------------------------------------------------------------------------------------------------
struct CovPlan {
    int threads, LDS, LDV, AS, CH;
    bool global_m;
    size_t smem, per_sample_ws;
};

static int cov_smem_limit() { return device_smem_limit(); }   // of the current device

static CovPlan cov_plan_fwd(int L, int N, int tsize, int do_sqrt) {
    CovPlan pl;
    pl.threads = N <= 64 ? 128 : 256;
    pl.LDS = N | 1;
    pl.LDV = round_up(N, 4);
    pl.AS = round_up(N, 4);
    pl.CH = max(1, min(max(L, 1), pl.threads * kPre / N));
    const int NP = round_up(N + 2, 32);
    const size_t vec = ((size_t)2 * pl.CH * pl.AS + 4 * (size_t)NP + 64) * tsize + 2 * (size_t)NP * sizeof(int);
    const size_t mat = ((size_t)round_up(N * pl.LDS, 4) + (do_sqrt ? (size_t)round_up(N * pl.LDV, 4) : 0)) * tsize;
    pl.global_m = false;
    pl.smem = vec + mat;
    pl.per_sample_ws = 0;
    if (pl.smem > (size_t)cov_smem_limit()) {
        pl.global_m = true;
        pl.smem = vec;
        pl.per_sample_ws = ((size_t)round_up(N * pl.LDS, 4) + (size_t)round_up(N * pl.LDV, 4)) * tsize;
    }
    return pl;
}

static CovPlan cov_plan_bwd(int L, int N, int tsize, int do_sqrt) {
    CovPlan pl;
    pl.threads = N <= 64 ? 128 : 256;
    pl.LDS = N | 1;
    pl.LDV = pl.LDS;
    pl.AS = round_up(N, 4);
    pl.CH = max(2, min(max(L, 2), pl.threads * kPre / N));   // >= 2 so that the staging area also holds the 4 x N d_x tile
    const int NP = round_up(N + 2, 32);
    const size_t vec = ((size_t)2 * pl.CH * pl.AS + 2 * (size_t)NP + 64) * tsize;
    const size_t msz = (size_t)round_up(N * pl.LDS, 4);
    const size_t mat = (size_t)(do_sqrt ? 3 : 2) * msz * tsize;
    pl.global_m = false;
    pl.smem = vec + mat;
    pl.per_sample_ws = 0;
    if (pl.smem > (size_t)cov_smem_limit()) {
        pl.global_m = true;
        pl.smem = vec;
        pl.per_sample_ws = 3 * msz * tsize;
    }
    return pl;
}

// eig_block.cu: float32 forward for matrices beyond shared memory (tensor-core Gram + blocked Jacobi)
size_t eig_block_workspace_bytes(long long B, int N);
int covariance_fwd_large_f32(const float *x, const float *coef, int strategy, int do_sqrt, long long B, int L, int N, float *out,
                             float *eigvec, float *eigval, int32_t *sweeps, void *ws, size_t ws_bytes, cudaStream_t st);

size_t covariance_workspace_bytes(long long B, int L, int N, int tsize, int do_sqrt) {
    const CovPlan pf = cov_plan_fwd(L, N, tsize, do_sqrt), pb = cov_plan_bwd(L, N, tsize, do_sqrt);
    size_t fwd = (size_t)B * pf.per_sample_ws;
    if (pf.global_m && tsize == 4) fwd = eig_block_workspace_bytes(B, N);
    return max(fwd, (size_t)B * pb.per_sample_ws);
}

template <typename T>
int covariance_fwd_t(const void *x, const void *coef, int strategy, int do_sqrt, long long B, int L, int N, void *out,
                     void *eigvec, void *eigval, int32_t *sweeps, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (!do_sqrt && cov_plain_fits(L, N, sizeof(T))) {
        CovParams<T> q;
        memset(&q, 0, sizeof(q));
        q.x = (const T *)x; q.coef = (const T *)coef; q.strategy = strategy; q.B = B; q.L = L; q.N = N;
        q.AS = round_up(N, 4); q.out = (T *)out;
        const size_t smem = cov_plain_smem(L, N, sizeof(T));
        static DeviceOnce configured;     // per dtype instance and per device: opt in to the plain kernel's ceiling once
        const int dev = current_device();
        if (smem > 48 * 1024 && configured.needs(dev)) {
            if (cudaFuncSetAttribute(covariance_plain_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 75 * 1024) != cudaSuccess)
                return check_launch("cudaFuncSetAttribute(covariance_plain_kernel)");
            configured.mark(dev);
        }
        covariance_plain_kernel<T><<<(unsigned)B, kPlainThreads, smem, st>>>(q);
        return check_launch("covariance_plain_kernel");
    }
    const CovPlan pl = cov_plan_fwd(L, N, sizeof(T), do_sqrt);
    if (pl.global_m && sizeof(T) == 4)     // beyond shared memory: tensor-core Gram + blocked Jacobi (eig_block.cu)
        return covariance_fwd_large_f32((const float *)x, (const float *)coef, strategy, do_sqrt, B, L, N, (float *)out, (float *)eigvec,
                                        (float *)eigval, sweeps, ws, ws_bytes, st);
    if (pl.global_m && (ws_bytes < (size_t)B * pl.per_sample_ws || !ws)) { set_error("ddw_covariance_fwd: workspace too small"); return DDW_ERR_WORKSPACE; }
    if (pl.smem > (size_t)cov_smem_limit()) { set_error("ddw_covariance_fwd: n_assets=%d needs %zu B of shared memory", N, pl.smem); return DDW_ERR_UNSUPPORTED; }
    CovParams<T> p;
    memset(&p, 0, sizeof(p));
    p.x = (const T *)x; p.coef = (const T *)coef; p.strategy = strategy; p.do_sqrt = do_sqrt; p.B = B; p.L = L; p.N = N;
    p.LDS = pl.LDS; p.LDV = pl.LDV; p.AS = pl.AS; p.CH = pl.CH;
    p.out = (T *)out; p.eigvec = (T *)eigvec; p.eigval = (T *)eigval; p.sweeps = sweeps; p.ws = (T *)ws;
#define DDW_COV_LAUNCH(KERNEL, TH, GL)                                                                      \
    do {                                                                                                    \
        auto k = KERNEL<T, TH, GL>;                                                                         \
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess) \
            return check_launch("cudaFuncSetAttribute(" #KERNEL ")");                                        \
        k<<<(unsigned)B, TH, pl.smem, st>>>(p);                                                             \
        return check_launch(#KERNEL);                                                                       \
    } while (0)
    if (pl.threads == 128) { if (pl.global_m) DDW_COV_LAUNCH(covariance_fwd_kernel, 128, true); else DDW_COV_LAUNCH(covariance_fwd_kernel, 128, false); }
    if (pl.global_m) DDW_COV_LAUNCH(covariance_fwd_kernel, 256, true);
    DDW_COV_LAUNCH(covariance_fwd_kernel, 256, false);
}

template <typename T>
int covariance_bwd_t(const void *grad_out, const void *x, const void *coef, int strategy, int do_sqrt,
                     const void *eigvec, const void *eigval, long long B, int L, int N, void *d_x, void *d_coef,
                     void *ws, size_t ws_bytes, cudaStream_t st) {
    const CovPlan pl = cov_plan_bwd(L, N, sizeof(T), do_sqrt);
    if (pl.global_m && (ws_bytes < (size_t)B * pl.per_sample_ws || !ws)) { set_error("ddw_covariance_bwd: workspace too small"); return DDW_ERR_WORKSPACE; }
    if (pl.smem > (size_t)cov_smem_limit()) { set_error("ddw_covariance_bwd: n_assets=%d needs %zu B of shared memory", N, pl.smem); return DDW_ERR_UNSUPPORTED; }
    CovParams<T> p;
    memset(&p, 0, sizeof(p));
    p.x = (const T *)x; p.coef = (const T *)coef; p.grad_out = (const T *)grad_out; p.eigvec_in = (const T *)eigvec;
    p.eigval_in = (const T *)eigval; p.strategy = strategy; p.do_sqrt = do_sqrt; p.B = B; p.L = L; p.N = N;
    p.LDS = pl.LDS; p.LDV = pl.LDV; p.AS = pl.AS; p.CH = pl.CH; p.d_x = (T *)d_x; p.d_coef = (T *)d_coef; p.ws = (T *)ws;
    if (pl.threads == 128) { if (pl.global_m) DDW_COV_LAUNCH(covariance_bwd_kernel, 128, true); else DDW_COV_LAUNCH(covariance_bwd_kernel, 128, false); }
    if (pl.global_m) DDW_COV_LAUNCH(covariance_bwd_kernel, 256, true);
    DDW_COV_LAUNCH(covariance_bwd_kernel, 256, false);
#undef DDW_COV_LAUNCH
}

template int covariance_fwd_t<float>(const void *, const void *, int, int, long long, int, int, void *, void *, void *,
                                     int32_t *, void *, size_t, cudaStream_t);
template int covariance_fwd_t<double>(const void *, const void *, int, int, long long, int, int, void *, void *,
                                      void *, int32_t *, void *, size_t, cudaStream_t);
template int covariance_bwd_t<float>(const void *, const void *, const void *, int, int, const void *, const void *,
                                     long long, int, int, void *, void *, void *, size_t, cudaStream_t);
template int covariance_bwd_t<double>(const void *, const void *, const void *, int, int, const void *, const void *,
                                      long long, int, int, void *, void *, void *, size_t, cudaStream_t);

}  // namespace ddw
