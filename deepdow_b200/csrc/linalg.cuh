// linalg.cuh -- CTA-level dense linear algebra shared by the solver and covariance kernels
#pragma once
#include "common.cuh"

namespace ddw {

constexpr int kPre = 8;   // staged elements per thread per chunk (register prefetch depth)

// ------------------------------------------------------------------------------------------------
// Gram of a K x N row-major matrix: out(i,j) = sum_k A[k,ci] A[k,cj] (+ diag_add), lower triangle.
// A[k,i] = Cb[k,i] * gamma[(gflat0 + k N + i) % B] when gamma != nullptr (the NumericalMarkowitz
// tiling), minus mean[i] when mean != nullptr (CovarianceMatrix centring).
// Rows of A are streamed through a double-buffered shared staging area; with `pos` != nullptr only
// the columns with pos[i] >= 0 are kept (compacted to column pos[i]).
// ------------------------------------------------------------------------------------------------
template <typename T, int kThreads, bool kColMajorOut>
__device__ __forceinline__ void gram_lower(const T *__restrict__ Cb, const T *__restrict__ gamma,
                                           long long gflat0, long long B, int K, int N, const T *mean,
                                           const int *pos, int ncols, T *As, int AS, int CH, T *out, int ld,
                                           T diag_add) {
    const int tid = threadIdx.x;
    const int NT = (ncols + 3) >> 2;
    const int ntiles = NT * (NT + 1) / 2;
    T pre[kPre];
    const int dk = kThreads / N, di = kThreads % N;
    const unsigned gstep = gamma ? (unsigned)(kThreads % B) : 0u;

    auto fetch = [&](int row0) {
        const int rows = min(CH, K - row0);
        const int cnt = rows * N;
        const long long base = (long long)row0 * N;
        unsigned gi = gamma ? (unsigned)((gflat0 + base + tid) % B) : 0u;
#pragma unroll
        for (int j = 0; j < kPre; ++j) {
            const int e = tid + j * kThreads;
            if (e < cnt) pre[j] = gamma ? Cb[base + e] * gamma[gi] : Cb[base + e];
            gi += gstep;
            if (gi >= (unsigned)B) gi -= (unsigned)B;
        }
    };
    auto stash = [&](T *buf, int row0) {
        const int rows = min(CH, K - row0);
        const int cnt = rows * N;
        int k = tid / N, i = tid - k * N;
#pragma unroll
        for (int j = 0; j < kPre; ++j) {
            const int e = tid + j * kThreads;
            if (e < cnt) {
                const int col = pos ? pos[i] : i;
                if (col >= 0) buf[k * AS + col] = mean ? pre[j] - mean[i] : pre[j];
            }
            i += di; k += dk;
            if (i >= N) { i -= N; ++k; }
        }
    };

    for (int e = tid; e < 2 * CH * AS; e += kThreads) As[e] = T(0);
    fetch(0);
    __syncthreads();
    int buf = 0;
    for (int row0 = 0; row0 < K; row0 += CH, buf ^= 1) {
        T *Ab = As + buf * CH * AS;
        stash(Ab, row0);
        __syncthreads();
        if (row0 + CH < K) fetch(row0 + CH);
        const int rows = min(CH, K - row0);
        const bool first = row0 == 0;
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
            const T *pa = Ab + 4 * ti, *pb = Ab + 4 * tj;
            for (int kk = 0; kk < rows; ++kk) {
                const Vec4<T> va = ld4(pa + kk * AS), vb = ld4(pb + kk * AS);
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = fma(va.v[x], vb.v[y], acc[x][y]);
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const int i = 4 * ti + x, j = 4 * tj + y;
                    if (i < ncols && j <= i) {
                        T *o = out + (kColMajorOut ? (j * ld + i) : (i * ld + j));
                        *o = first ? acc[x][y] + (i == j ? diag_add : T(0)) : *o + acc[x][y];
                    }
                }
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// In-place unscaled LDL' of the n x n lower triangle stored column-major at Lb (element (i,k) at
// Lb[k*ld + i]) with `nrhs` extra rows n .. n+nrhs-1 that undergo the same elimination (so row
// n+q ends up holding D * L^-1 * rhs_q).  dinv[k] = 1 / pivot_k.  One barrier per column.
// Returns nonzero when a pivot had to be lifted.
// ------------------------------------------------------------------------------------------------
template <typename T, int kThreads>
__device__ __forceinline__ int factor_ldl(T *Lb, int ld, int n, int nrhs, T *dinv, T dmin) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kThreads / 32;
    int lifted = 0;
    const int rows = n + nrhs;
    for (int k = 0; k < n; ++k) {
        const T *colk = Lb + k * ld;
        T d = colk[k];
        if (!(d > dmin)) { d = dmin; lifted = 1; }
        const T invd = T(1) / d;
        if (threadIdx.x == 0) dinv[k] = invd;
        for (int j = k + 1 + warp; j < n; j += kWarps) {
            const T cj = colk[j] * invd;
            T *colj = Lb + j * ld;
            for (int i = j + lane; i < rows; i += 32) colj[i] = fma(-colk[i], cj, colj[i]);
        }
        __syncthreads();
    }
    return lifted;
}

// Back substitution L' x = z for the factor above, executed by ONE warp with z/x held in registers:
// element k lives in lane k%32, slot k/32.  On return zt holds x.
template <typename T, int NREG>
__device__ __forceinline__ void backsub_warp(const T *Lb, int ld, int n, const T *dinv, T (&zt)[NREG]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int mm = NREG - 1; mm >= 0; --mm) {
        if (mm * 32 >= n) continue;
        const int ktop = min(n - 1 - mm * 32, 31);
        for (int kk = ktop; kk >= 0; --kk) {
            const int k = mm * 32 + kk;
            const T xk = __shfl_sync(kFullMask, zt[mm], kk) * dinv[k];
            if (lane == kk) zt[mm] = xk;
#pragma unroll
            for (int m2 = 0; m2 < NREG; ++m2) {
                if (m2 <= mm) {
                    const int j = lane + 32 * m2;
                    if (j < k) zt[m2] = fma(-Lb[j * ld + k], xk, zt[m2]);
                }
            }
        }
    }
}

// Forward substitution t = L^-1 rhs (same storage), one warp, registers; on return zt holds D*t,
// i.e. exactly what an eliminated right-hand-side row would hold.
template <typename T, int NREG>
__device__ __forceinline__ void fwdsub_warp(const T *Lb, int ld, int n, const T *dinv, T (&zt)[NREG]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int mm = 0; mm < NREG; ++mm) {
        if (mm * 32 >= n) continue;
        const int ktop = min(n - mm * 32, 32);
        for (int kk = 0; kk < ktop; ++kk) {
            const int k = mm * 32 + kk;
            const T tk = __shfl_sync(kFullMask, zt[mm], kk) * dinv[k];
#pragma unroll
            for (int m2 = 0; m2 < NREG; ++m2) {
                if (m2 >= mm) {
                    const int i = lane + 32 * m2;
                    if (i > k && i < n) zt[m2] = fma(-Lb[k * ld + i], tk, zt[m2]);
                }
            }
        }
    }
}

}  // namespace ddw
