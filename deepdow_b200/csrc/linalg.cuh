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
// Same contract as gram_lower, for matrices small enough that every thread can keep its (at most
// MAXT) 4x4 output tiles in registers across the whole K loop: the shared staging area is only
// read, the result is written once at the end.  Staging uses 16-byte loads/stores when the chunk is
// a flat copy (no compaction, no centring, N % 4 == 0).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st4(float *p, const Vec4<float> &v) {
    *reinterpret_cast<float4 *>(p) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
}
__device__ __forceinline__ void st4(double *p, const Vec4<double> &v) {
    *reinterpret_cast<double2 *>(p) = make_double2(v.v[0], v.v[1]);
    *reinterpret_cast<double2 *>(p + 2) = make_double2(v.v[2], v.v[3]);
}

template <typename T, int kThreads, int MAXT, bool kColMajorOut>
__device__ __forceinline__ void gram_lower_reg(const T *__restrict__ Cb, const T *__restrict__ gamma,
                                               long long gflat0, long long B, int K, int N, const T *mean,
                                               const int *pos, int ncols, T *As, int AS, int CH, T *out, int ld,
                                               T diag_add) {
    const int tid = threadIdx.x;
    const int NT = (ncols + 3) >> 2;
    const int ntiles = NT * (NT + 1) / 2;
    const bool flat = (AS == N) && pos == nullptr && mean == nullptr && (N & 3) == 0;
    T pre[kPre];
    const int dk = kThreads / N, di = kThreads % N;
    const unsigned uB = (unsigned)B;
    const unsigned gstep = gamma ? (unsigned)(kThreads % B) : 0u;
    const unsigned gstep4 = gamma ? (unsigned)((4 * kThreads) % B) : 0u;

    int offa[MAXT], offb[MAXT];
    T acc[MAXT][4][4];
#pragma unroll
    for (int m = 0; m < MAXT; ++m) {
        const int t = tid + m * kThreads;
        offa[m] = -1; offb[m] = 0;
        if (t < ntiles) {
            int ti = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
            while (ti * (ti + 1) / 2 > t) --ti;
            while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
            offa[m] = 4 * ti; offb[m] = 4 * (t - ti * (ti + 1) / 2);
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[m][x][y] = T(0);
    }

    // fetch() only issues the global loads (matrix element and its gamma factor); the product is
    // formed in stash(), one chunk later, so the loads stay in flight during the tile FMAs
    T gpre[kPre];
    auto fetch = [&](int row0) {
        const int rows = min(CH, K - row0);
        const int cnt = rows * N;
        const long long base = (long long)row0 * N;
        if (flat) {
            unsigned gi = gamma ? (unsigned)((gflat0 + base + 4 * tid) % B) : 0u;
#pragma unroll
            for (int j = 0; j < kPre / 4; ++j) {
                const int e = 4 * (tid + j * kThreads);
                if (e < cnt) {
                    const Vec4<T> v = ld4(Cb + base + e);
#pragma unroll
                    for (int q = 0; q < 4; ++q) pre[4 * j + q] = v.v[q];
                    if (gamma) {
                        if ((uB & 3u) == 0u) {
                            const Vec4<T> g4 = ld4(gamma + gi);
#pragma unroll
                            for (int q = 0; q < 4; ++q) gpre[4 * j + q] = g4.v[q];
                        } else {
                            unsigned g = gi;
#pragma unroll
                            for (int q = 0; q < 4; ++q) { gpre[4 * j + q] = gamma[g]; if (++g >= uB) g = 0; }
                        }
                    }
                }
                gi += gstep4;
                if (gi >= uB) gi -= uB;
            }
        } else {
            unsigned gi = gamma ? (unsigned)((gflat0 + base + tid) % B) : 0u;
#pragma unroll
            for (int j = 0; j < kPre; ++j) {
                const int e = tid + j * kThreads;
                if (e < cnt) { pre[j] = Cb[base + e]; if (gamma) gpre[j] = gamma[gi]; }
                gi += gstep;
                if (gi >= uB) gi -= uB;
            }
        }
    };
    auto stash = [&](T *buf, int row0) {
        const int rows = min(CH, K - row0);
        const int cnt = rows * N;
        if (flat) {
#pragma unroll
            for (int j = 0; j < kPre / 4; ++j) {
                const int e = 4 * (tid + j * kThreads);
                if (e < cnt) {
                    Vec4<T> v;
#pragma unroll
                    for (int q = 0; q < 4; ++q) v.v[q] = gamma ? pre[4 * j + q] * gpre[4 * j + q] : pre[4 * j + q];
                    st4(buf + e, v);
                }
            }
        } else {
            int k = tid / N, i = tid - k * N;
#pragma unroll
            for (int j = 0; j < kPre; ++j) {
                const int e = tid + j * kThreads;
                if (e < cnt) {
                    const int col = pos ? pos[i] : i;
                    T val = gamma ? pre[j] * gpre[j] : pre[j];
                    if (mean) val -= mean[i];
                    if (col >= 0) buf[k * AS + col] = val;
                }
                i += di; k += dk;
                if (i >= N) { i -= N; ++k; }
            }
        }
    };

    if (!flat) for (int e = tid; e < 2 * CH * AS; e += kThreads) As[e] = T(0);
    fetch(0);
    __syncthreads();
    int buf = 0;
    for (int row0 = 0; row0 < K; row0 += CH, buf ^= 1) {
        T *Ab = As + buf * CH * AS;
        stash(Ab, row0);
        __syncthreads();
        if (row0 + CH < K) fetch(row0 + CH);
        const int rows = min(CH, K - row0);
#pragma unroll 2
        for (int kk = 0; kk < rows; ++kk) {
            const T *rowp = Ab + kk * AS;
#pragma unroll
            for (int m = 0; m < MAXT; ++m) {
                if (offa[m] >= 0) {
                    const Vec4<T> va = ld4(rowp + offa[m]), vb = ld4(rowp + offb[m]);
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) acc[m][x][y] = fma(va.v[x], vb.v[y], acc[m][x][y]);
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MAXT; ++m) {
        if (offa[m] >= 0) {
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const int i = offa[m] + x, j = offb[m] + y;
                    if (i < ncols && j <= i)
                        out[kColMajorOut ? (j * ld + i) : (i * ld + j)] = acc[m][x][y] + (i == j ? diag_add : T(0));
                }
        }
    }
    __syncthreads();
}

// number of 4x4 lower-triangle tiles per thread needed by gram_lower_reg
__host__ __device__ __forceinline__ int gram_tiles_per_thread(int n, int threads) {
    const int nt = (n + 3) / 4;
    return (nt * (nt + 1) / 2 + threads - 1) / threads;
}

// ------------------------------------------------------------------------------------------------
// In-place unscaled LDL' of the n x n lower triangle stored column-major at Lb (element (i,k) at
// Lb[k*ld + i]) with `nrhs` extra rows n .. n+nrhs-1 that undergo the same elimination (so row
// n+q ends up holding D * L^-1 * rhs_q).  dinv[k] = 1 / pivot_k.
// Returns nonzero when a pivot had to be lifted.
// ------------------------------------------------------------------------------------------------
// Blocked right-looking form, panels of kLdlNb columns, two block barriers per panel:
//   panel     every thread factors the kLdlNb x kLdlNb diagonal block redundantly in registers (85 FMAs), then eliminates
//             its own row(s) below the block against it -- no barrier inside the panel;
//   trailing  warps take blocks of 4 columns, lanes take rows: the panel's entries of a row are loaded once and used for
//             all 4 columns, the 4 x kLdlNb multipliers live in registers (1.5 instructions per FMA).
// The column-at-a-time form this replaces (one barrier and one short row loop per (k, j) pair) spent 19 instructions per
// FMA at n = 86: 80% of the dense forward kernel's instructions (ncu, profiles/r01_final/dense_regime.txt).
constexpr int kLdlNb = 8;

template <typename T, int kThreads>
__device__ __forceinline__ int factor_ldl(T *Lb, int ld, int n, int nrhs, T *dinv, T dmin) {
    constexpr int NB = kLdlNb, JB = 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;
    int lifted = 0;
    const int rows = n + nrhs;
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int kb = min(NB, n - k0);
        // ---- panel: diagonal block in registers (lower triangle, element (a, c) of the block at Bd[a][c], c <= a)
        T Bd[NB][NB], inv[NB];
        const bool fullp = kb == NB;         // full panel: no bounds tests anywhere below (they cost as much as the arithmetic)
        if (fullp) {
#pragma unroll
            for (int c = 0; c < NB; ++c)
#pragma unroll
                for (int a2 = c; a2 < NB; ++a2) Bd[a2][c] = Lb[(k0 + c) * ld + k0 + a2];
        } else {
#pragma unroll
            for (int c = 0; c < NB; ++c)
#pragma unroll
                for (int a2 = c; a2 < NB; ++a2) Bd[a2][c] = (a2 < kb) ? Lb[(k0 + c) * ld + k0 + a2] : T(0);
        }
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            T d = Bd[c][c];
            if (c < kb && !(d > dmin)) { d = dmin; lifted = 1; }
            inv[c] = c < kb ? recipFastT(d) : T(0);     // d > dmin > 0: the hardware reciprocal + one Newton step is within 1 ulp
#pragma unroll
            for (int j = c + 1; j < NB; ++j) {
                const T cj = Bd[j][c] * inv[c];
#pragma unroll
                for (int a2 = j; a2 < NB; ++a2) Bd[a2][j] = fma(-Bd[a2][c], cj, Bd[a2][j]);
            }
        }
#pragma unroll
        for (int c = 0; c < NB; ++c)
            if (tid == c && c < kb) dinv[k0 + c] = inv[c];
        // rows below the block: one row per thread and pass, eliminated against the block held in registers
        for (int i = k0 + tid; i < rows; i += kThreads) {
            const int t = i - k0;
            if (t >= kb) {
                T v[NB];
                T *pr = Lb + k0 * ld + i;
                if (fullp) {
#pragma unroll
                    for (int c = 0; c < NB; ++c) v[c] = pr[c * ld];
                } else {
#pragma unroll
                    for (int c = 0; c < NB; ++c) v[c] = c < kb ? pr[c * ld] : T(0);
                }
#pragma unroll
                for (int c = 0; c < NB; ++c) {
                    const T vc = v[c] * inv[c];
#pragma unroll
                    for (int j = c + 1; j < NB; ++j) v[j] = fma(-vc, Bd[j][c], v[j]);
                }
                if (fullp) {
#pragma unroll
                    for (int c = 1; c < NB; ++c) pr[c * ld] = v[c];
                } else {
#pragma unroll
                    for (int c = 1; c < NB; ++c)
                        if (c < kb) pr[c * ld] = v[c];
                }
            }
        }
        __syncthreads();
        // the factored rows of the diagonal block go back only now: during the panel every thread was still reading the
        // block, and nothing below reads these entries (the trailing update takes rows >= k0 + kb of the panel columns)
        if (tid < kb) {
#pragma unroll
            for (int a2 = 0; a2 < NB; ++a2)
                if (a2 == tid) {
#pragma unroll
                    for (int c = 0; c <= a2; ++c) Lb[(k0 + c) * ld + k0 + tid] = Bd[a2][c];
                }
        }
        // ---- trailing update: col_j[i] -= sum_c col_c[i] * col_c[j] / d_c, c over the panel, j >= k0 + kb, i >= j
        const int j0 = k0 + kb;
        for (int jb = j0 + JB * warp; jb < n; jb += JB * kWarps) {
            T cf[JB][NB];
            if (kb == NB && jb + JB <= n) {
                // full panel, full column block: no bounds tests; only the rows inside the block's own triangle are masked
#pragma unroll
                for (int q = 0; q < JB; ++q)
#pragma unroll
                    for (int c = 0; c < NB; ++c) cf[q][c] = Lb[(k0 + c) * ld + jb + q] * inv[c];
                for (int i = jb + lane; i < rows; i += 32) {
                    T av[NB];
#pragma unroll
                    for (int c = 0; c < NB; ++c) av[c] = Lb[(k0 + c) * ld + i];
                    T *pe = Lb + jb * ld + i;
#pragma unroll
                    for (int q = 0; q < JB; ++q) {
                        if (i >= jb + q) {
                            T acc = pe[q * ld];
#pragma unroll
                            for (int c = 0; c < NB; ++c) acc = fma(-av[c], cf[q][c], acc);
                            pe[q * ld] = acc;
                        }
                    }
                }
                continue;
            }
#pragma unroll
            for (int q = 0; q < JB; ++q)
#pragma unroll
                for (int c = 0; c < NB; ++c) cf[q][c] = (jb + q < n && c < kb) ? Lb[(k0 + c) * ld + jb + q] * inv[c] : T(0);
            for (int i = jb + lane; i < rows; i += 32) {
                T av[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c) av[c] = c < kb ? Lb[(k0 + c) * ld + i] : T(0);
#pragma unroll
                for (int q = 0; q < JB; ++q) {
                    if (jb + q < n && i >= jb + q) {
                        T *pe = Lb + (jb + q) * ld + i;
                        T acc = *pe;
#pragma unroll
                        for (int c = 0; c < NB; ++c) acc = fma(-av[c], cf[q][c], acc);
                        *pe = acc;
                    }
                }
            }
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

// Two right-hand sides through the same forward substitution: the factor entries are loaded once and the two dependent
// chains (shuffle -> multiply -> FMA) interleave, which is what a lone warp needs to hide their latency.
template <typename T, int NREG>
__device__ __forceinline__ void fwdsub2_warp(const T *Lb, int ld, int n, const T *dinv, T (&z1)[NREG], T (&z2)[NREG]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int mm = 0; mm < NREG; ++mm) {
        if (mm * 32 >= n) continue;
        const int ktop = min(n - mm * 32, 32);
        for (int kk = 0; kk < ktop; ++kk) {
            const int k = mm * 32 + kk;
            const T di = dinv[k];
            const T t1 = __shfl_sync(kFullMask, z1[mm], kk) * di, t2 = __shfl_sync(kFullMask, z2[mm], kk) * di;
#pragma unroll
            for (int m2 = 0; m2 < NREG; ++m2) {
                if (m2 >= mm) {
                    const int i = lane + 32 * m2;
                    if (i > k && i < n) {
                        const T l = Lb[k * ld + i];
                        z1[m2] = fma(-l, t1, z1[m2]);
                        z2[m2] = fma(-l, t2, z2[m2]);
                    }
                }
            }
        }
    }
}

}  // namespace ddw
