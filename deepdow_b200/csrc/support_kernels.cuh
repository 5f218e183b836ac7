// support_kernels.cuh -- building blocks of the sparse-support NumericalMarkowitz kernels (forward and backward):
// A (gamma-tiled covmat_sqrt) sits in shared memory, row-major N x N, and everything is computed from gathered
// columns of it -- column passes and 4x4 Gram tiles of selected columns.  (The reduced KKT solve is the four-warp
// register factorisation of markowitz_common.cuh; the one-warp forms that were tried are recorded in
// profiles/r01_final/phase_timing.txt and in the history of this file.)
#pragma once
#include "common.cuh"
#include "linalg.cuh"

namespace ddw {

constexpr int kPartElems = 512;    // partial sums of the column passes (KS slices x N columns)
constexpr int kGramLanes = 8;      // lanes that share one 4x4 Gram tile (rows k = q, q+8, ...)

// A -> shared memory by one bulk TMA copy when size and address allow it, else a plain cooperative copy.
// Call from all threads; returns after the data has landed (no block barrier: the caller syncs before
// other threads' elements are read).
template <typename T, int kThreads>
__device__ __forceinline__ bool stage_matrix_begin(T *As, const T *Cb, int N, unsigned long long *mbar) {
    const unsigned bytes = (unsigned)(N * N * sizeof(T));
    const bool use_tma = (bytes & 15u) == 0u && (reinterpret_cast<unsigned long long>(Cb) & 15ull) == 0ull;
    if (use_tma) {
        if (threadIdx.x == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
        __syncthreads();
        if (threadIdx.x == 0) { mbar_expect_tx(mbar, bytes); tma_bulk_g2s(As, Cb, bytes, mbar); }
    }
    return use_tma;
}
template <typename T, int kThreads>
__device__ __forceinline__ void stage_matrix_end(T *As, const T *Cb, int N, unsigned long long *mbar, bool use_tma) {
    if (use_tma) mbar_wait(mbar, 0);
    else for (int e = threadIdx.x; e < N * N; e += kThreads) As[e] = Cb[e];
}

// gamma tiling (allocate.py:268-270) applied in place: As[e] *= gamma[(gflat0 + e) % Bmod].  Every thread touches
// only elements it has itself waited for; the caller syncs afterwards.
template <typename T, int kThreads>
__device__ __forceinline__ void apply_gamma_tiling(T *As, int N, const T *__restrict__ gamma, long long gflat0, long long Bmod) {
    const int tid = threadIdx.x;
    const unsigned uB = (unsigned)Bmod;
    if ((N & 3) == 0 && (uB & 3u) == 0u) {
        unsigned gi = (unsigned)((gflat0 + 4 * tid) % Bmod);
        const unsigned gstep = (unsigned)((4 * kThreads) % Bmod);
        for (int e = 4 * tid; e < N * N; e += 4 * kThreads) {
            Vec4<T> a = ld4(As + e);
            const Vec4<T> g4 = ld4(gamma + gi);
#pragma unroll
            for (int q = 0; q < 4; ++q) a.v[q] *= g4.v[q];
            st4(As + e, a);
            gi += gstep; if (gi >= uB) gi -= uB;
        }
    } else {
        unsigned gi = (unsigned)((gflat0 + tid) % Bmod);
        const unsigned gstep = (unsigned)(kThreads % Bmod);
        for (int e = tid; e < N * N; e += kThreads) {
            As[e] *= gamma[gi];
            gi += gstep; if (gi >= uB) gi -= uB;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Column pass over A: sum_k As[k,i] * f(k), with f(k) = t[k] or, when kSquare, As[k,i] itself (column norms).
// N % 4 == 0: thread (group of 4 columns, row slice) accumulates with 16-byte loads and writes its partial sums
// to part[slice*N + i]; the caller syncs and adds the `slices` rows (column_pass_sum).  Otherwise one thread per
// column, a single slice.  Returns the number of slices.
// ------------------------------------------------------------------------------------------------
template <typename T, int kThreads, bool kSquare>
__device__ __forceinline__ int column_pass(const T *As, int N, const T *t, T *part) {
    const int tid = threadIdx.x;
    if ((N & 3) == 0) {
        const int NG = N >> 2;
        int KS = kThreads / NG;
        KS = min(KS, min(8, kPartElems / N));
        const int rows = (N + KS - 1) / KS;
        const int g = tid % NG, q = tid / NG;
        if (q < KS) {
            const int k0 = q * rows, k1 = min(N, k0 + rows);
            T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
            const T *col = As + 4 * g;
#pragma unroll 4
            for (int k = k0; k < k1; ++k) {
                const Vec4<T> a = ld4(col + k * N);
                if (kSquare) {
                    a0 = fma(a.v[0], a.v[0], a0); a1 = fma(a.v[1], a.v[1], a1);
                    a2 = fma(a.v[2], a.v[2], a2); a3 = fma(a.v[3], a.v[3], a3);
                } else {
                    const T tk = t[k];
                    a0 = fma(a.v[0], tk, a0); a1 = fma(a.v[1], tk, a1);
                    a2 = fma(a.v[2], tk, a2); a3 = fma(a.v[3], tk, a3);
                }
            }
            Vec4<T> o; o.v[0] = a0; o.v[1] = a1; o.v[2] = a2; o.v[3] = a3;
            st4(part + q * N + 4 * g, o);
        }
        return KS;
    }
    for (int i = tid; i < N; i += kThreads) {
        T a0 = T(0), a1 = T(0);
        int k = 0;
        for (; k + 1 < N; k += 2) {
            const T x0 = As[k * N + i], x1 = As[(k + 1) * N + i];
            a0 = fma(x0, kSquare ? x0 : t[k], a0);
            a1 = fma(x1, kSquare ? x1 : t[k + 1], a1);
        }
        if (k < N) { const T x0 = As[k * N + i]; a0 = fma(x0, kSquare ? x0 : t[k], a0); }
        part[i] = a0 + a1;
    }
    return 1;
}

template <typename T> __device__ __forceinline__ T column_pass_sum(const T *part, int N, int slices, int i) {
    T acc = part[i];
    for (int q = 1; q < slices; ++q) acc += part[q * N + i];
    return acc;
}

// ------------------------------------------------------------------------------------------------
// Gram entries of columns cols[nc0 .. nc1-1] against columns cols[0 .. nc1-1]: G(s,c) = a_s . a_c for c <= s,
// written to out[s*rs + c*cs].  Diagonal entries take qd[cols[s]] when qd != nullptr (precomputed column norms,
// which then already carry any diagonal shift), else the computed norm + diag_add.
// Work items are 4x4 tiles, each shared by kGramLanes lanes that take rows k = q, q + 8, ... and reduce through a
// 3-stage butterfly that halves the number of live sums per stage.  All threads of the CTA must call it.
// ------------------------------------------------------------------------------------------------
template <typename T, int kThreads>
__device__ __forceinline__ void gram_columns(const T *As, int N, const int *cols, const T *qd, T diag_add, int nc0,
                                             int nc1, T *out, int rs, int cs) {
    const int tid = threadIdx.x, q = tid & (kGramLanes - 1);
    const int RG = (nc1 - nc0 + 3) >> 2;          // row groups (new columns)
    int total = 0;                                 // tiles: row group rg covers column groups 0 .. (smax >> 2)
    for (int rg = 0; rg < RG; ++rg) total += ((min(nc1 - 1, nc0 + 4 * rg + 3)) >> 2) + 1;
    // every lane of a warp runs the same number of rounds (the butterfly shuffles need all of them)
    const int rounds = (total + kThreads / kGramLanes - 1) / (kThreads / kGramLanes);
    for (int rnd = 0; rnd < rounds; ++rnd) {
        const int item = rnd * (kThreads / kGramLanes) + tid / kGramLanes;
        const bool live = item < total;
        int rg = 0, rem = live ? item : 0;
        for (;; ++rg) {
            const int cnt = ((min(nc1 - 1, nc0 + 4 * rg + 3)) >> 2) + 1;
            if (rem < cnt) break;
            rem -= cnt;
        }
        const int s0 = nc0 + 4 * rg, c0 = 4 * rem;
        int ia[4], ic[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            ia[x] = cols[min(s0 + x, nc1 - 1)];
            ic[x] = cols[min(c0 + x, nc1 - 1)];
        }
        T acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = T(0);
        if (live) {
#pragma unroll 2
            for (int k = q; k < N; k += kGramLanes) {
                const T *row = As + k * N;
                T a[4], c[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) { a[x] = row[ia[x]]; c[x] = row[ic[x]]; }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], c[y], acc[x][y]);
            }
        }
        // butterfly: the partner q^4 takes over two of the four rows, q^2 one of the two, q^1 two of the four columns
        T h8[2][4];
        {
            const bool up = (q & 4) != 0;
#pragma unroll
            for (int x = 0; x < 2; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const T keep = up ? acc[2 + x][y] : acc[x][y], give = up ? acc[x][y] : acc[2 + x][y];
                    h8[x][y] = keep + __shfl_xor_sync(kFullMask, give, 4);
                }
        }
        T h4[4];
        {
            const bool up = (q & 2) != 0;
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const T keep = up ? h8[1][y] : h8[0][y], give = up ? h8[0][y] : h8[1][y];
                h4[y] = keep + __shfl_xor_sync(kFullMask, give, 2);
            }
        }
        T h2[2];
        {
            const bool up = (q & 1) != 0;
#pragma unroll
            for (int y = 0; y < 2; ++y) {
                const T keep = up ? h4[2 + y] : h4[y], give = up ? h4[y] : h4[2 + y];
                h2[y] = keep + __shfl_xor_sync(kFullMask, give, 1);
            }
        }
        const int sx = s0 + ((q >> 2) & 1) * 2 + ((q >> 1) & 1);
#pragma unroll
        for (int y = 0; y < 2; ++y) {
            const int c = c0 + (q & 1) * 2 + y;
            if (live && sx < nc1 && c <= sx)
                out[sx * rs + c * cs] = (c == sx) ? (qd ? qd[cols[sx]] : h2[y] + diag_add) : h2[y];
        }
    }
}

}  // namespace ddw
