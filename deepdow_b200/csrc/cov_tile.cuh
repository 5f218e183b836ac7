// cov_tile.cuh -- the shared-memory tile code of CovarianceMatrix(sqrt=False): centred Gram + 1/(L-1) + shrinkage of ONE
// sample whose (L, N) lookback tile already sits in shared memory.  Shared by covariance_plain_kernel (covariance.cu) and
// by the fused covariance -> NumericalMarkowitz forward (markowitz_impl.cuh), so that both produce the same bits.
// Reference: deepdow/layers/misc.py:143-166.
#pragma once
#include "common.cuh"
#include "linalg.cuh"

namespace ddw {

constexpr int kPlainThreads = 128;
constexpr int kPlainParts = 4;     // row slices of the column-mean pass

// shared-memory elements of the tile: X (L, N) + Xc (L, XS) + S (N, N) + partial sums + reduction scratch
__host__ __device__ inline size_t cov_tile_elems(int L, int N) {
    const int XS = round_up(N, 4);
    return (size_t)round_up(L * N, 4) + (size_t)L * XS + (size_t)round_up(N * N, 4) + (size_t)kPlainParts * XS + 64;
}

// X (L, N) dense as in HBM -> S (N, N) dense, full symmetric.  Every thread of a kThreadsT-thread CTA calls it; the
// strides of the loops assume kThreadsT == kPlainThreads.  Ends with a barrier.
template <typename T>
__device__ __forceinline__ void cov_plain_tile(const T *X, T *Xc, T *S, T *part, T *red, int L, int N, int XS, int strategy, T c) {
    const int tid = threadIdx.x;
    // column means and centring with a fixed (column, row slice) per thread: the mean stays in a register and the loops
    // carry no index arithmetic.  P = row slices (as many as fit 128 threads; 1 and a column loop when XS > 128).
    {
        const int P = XS <= kPlainThreads ? min(kPlainParts, kPlainThreads / XS) : 1;
        const int q = XS <= kPlainThreads ? tid / XS : 0;
        const int i0 = XS <= kPlainThreads ? tid - q * XS : tid;
        const int PS = round_up(N, 4);
        const bool active = q < P;
        // pointer-stepped loops: with run-time strides the compiler spent ~25 instructions per element on index arithmetic
        const int rows = active ? (L - q + P - 1) / P : 0;       // rows q, q + P, ... of this thread
        const int xstep = P * N, cstep = P * XS;
        if (active)
            for (int i = i0; i < N; i += kPlainThreads) {
                const T *px = X + q * N + i;
                T a0 = T(0), a1 = T(0);
                int r = 0;
                for (; r + 1 < rows; r += 2) { a0 += px[0]; a1 += px[xstep]; px += 2 * xstep; }
                if (r < rows) a0 += px[0];
                part[q * PS + i] = a0 + a1;
            }
        __syncthreads();
        if (active) {
            const T invL = T(1) / (T)L;
            for (int i = i0; i < XS; i += kPlainThreads) {
                T *pc = Xc + q * XS + i;
                if (i < N) {
                    T m = T(0);
                    for (int r = 0; r < P; ++r) m += part[r * PS + i];
                    m *= invL;
                    const T *px = X + q * N + i;
#pragma unroll 4
                    for (int r = 0; r < rows; ++r) { *pc = *px - m; px += xstep; pc += cstep; }
                } else {
                    for (int r = 0; r < rows; ++r) { *pc = T(0); pc += cstep; }
                }
            }
        }
    }
    __syncthreads();
    const T fact = T(1) / (T)(L - 1);
    T tr = T(0);
    if (strategy == DDW_SHRINK_SCALED_IDENTITY) {   // mean of the sample variances (misc.py:157-161)
        T loc = T(0);
        for (int i = tid; i < N; i += kPlainThreads) {
            T a = T(0);
            for (int l = 0; l < L; ++l) a = fma(Xc[l * XS + i], Xc[l * XS + i], a);
            loc += a;
        }
        tr = block_sum<T, kPlainThreads>(loc, red) * fact / (T)N;
    }
    // S_ij = fact * acc * (c off the diagonal | shrinkage weight on it) + the target's diagonal (misc.py:149-166)
    const T woff = fact * c;
    const T wdiag = strategy == DDW_SHRINK_DIAGONAL ? fact * (c + (T(1) - c)) : woff;
    const T dadd = strategy == DDW_SHRINK_IDENTITY ? T(1) - c : (strategy == DDW_SHRINK_SCALED_IDENTITY ? (T(1) - c) * tr : T(0));
    const int NT = XS >> 2, ntiles = NT * (NT + 1) / 2;
    for (int t = tid; t < ntiles; t += kPlainThreads) {
        int ti = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
        while (ti * (ti + 1) / 2 > t) --ti;
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        const int tj = t - ti * (ti + 1) / 2;
        T acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = T(0);
        const T *pa = Xc + 4 * ti, *pb = Xc + 4 * tj;
#pragma unroll 4
        for (int l = 0; l < L; ++l) {
            const Vec4<T> a = ld4(pa), bb = ld4(pb);
            pa += XS; pb += XS;
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = fma(a.v[x], bb.v[y], acc[x][y]);
        }
        const int i0t = 4 * ti, j0t = 4 * tj;
        T *srow = S + i0t * N + j0t, *scol = S + j0t * N + i0t;      // the tile and its mirror image
        if (i0t + 3 < N && ti != tj) {                               // interior off-diagonal tile: no bounds, no diagonal
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const T v = acc[x][y] * woff;
                    srow[x * N + y] = v;
                    scol[y * N + x] = v;
                }
        } else {
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int i = i0t + x;
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const int j = j0t + y;
                    if (i < N && j < N) {
                        const T v = i == j ? fma(acc[x][y], wdiag, dadd) : acc[x][y] * woff;
                        srow[x * N + y] = v;
                        scol[y * N + x] = v;
                    }
                }
            }
        }
    }
    __syncthreads();
}

}  // namespace ddw
