// eig_block.cu -- CovarianceMatrix for large universes (float32, n_assets beyond one SM's shared memory) on tensor cores.
//
// Replaces, for n_assets > ~160, the per-sample work of CovarianceMatrix.forward (reference deepdow/layers/misc.py:107-119):
//   compute_covariance  (misc.py:143-166)  centred Gram X'X / (L - 1) + shrinkage  -> cov_gram_tc_kernel  (tcgen05, 3xTF32)
//   compute_sqrt        (misc.py:183-201)  SVD of the symmetric matrix             -> blocked Jacobi below
//   (v * sqrt(s)) v'    (misc.py:201)                                              -> recon_tc_kernel     (tcgen05, 3xTF32)
//
// Blocked two-sided Jacobi ("cyclic by blocks", blocks of 32 columns): a sweep is one round of neighbour pairs (2m, 2m+1) that
// rotates every index pair inside the 64-wide pair block, followed by nb - 1 round-robin rounds of disjoint block pairs (I, J)
// that rotate the 32 x 32 cross pairs.  Per round and sample:
//   bj_local_kernel      one CTA per block pair: the 64 x 64 pivot block in shared memory, 32 (63) parallel steps of 32 disjoint
//                        Rutishauser rotations, accumulated into U; writes E = U - I
//   bj_update_kernel<0>  Z[:, IJ] = (U' S[IJ, :])' and VT[IJ, :] = U' VT[IJ, :]   (128 x 64 x 64 tiles on tcgen05; Z lives in `W`)
//   bj_update_kernel<1>  S[IJ, :] = U' Z[IJ, :]                                   (= rows IJ of U' S U, which is symmetric)
// The updates run in "delta" form X + X E: the tensor core only forms the correction, which vanishes as the sweeps converge,
// and the base term is added back exactly in FP32 (numerical model: tools/models/block_jacobi_model.py).
// All samples of the batch advance together (grid = pairs x batch); a sample whose last sweep made no significant rotation is
// flagged done and its CTAs return at once.  Matrices live in the caller's workspace, padded to a multiple of 128.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace ddw {

constexpr int kBjB = 32;              // block width
constexpr int kBjW = 2 * kBjB;        // width of a block pair = K and N of the update GEMMs
constexpr int kBjMaxSweeps = 20;
constexpr float kBjSigRel = 1e-4f;    // a rotation is "significant" (another sweep is needed) above this relative size
constexpr int kLocalThreads = 256;      // 8 warps: the per-step dependent chains (LDS -> 4 FMAs -> STS, two barriers) need the extra warps

struct BjArgs {
    int N, NP, nb;                    // n_assets, padded size (multiple of 128), number of 32-wide blocks (NP / 32)
    long long B;
    float *S, *W, *VT, *ET;           // (B, NP, NP) x 3, (B, nb/2, 64, 64)
    float *mean, *rs, *trace;         // (B, NP), (B, NP), (B)
    int *state;                       // (B, 4): [0] significant rotations of the running sweep, [1] done, [2] sweeps used,
                                      //         [3] max |diag| of the input as float bits
};

__host__ __device__ inline int bj_padded(int N) { return (N + 127) / 128 * 128; }

static size_t bj_sample_bytes(int N) {
    const size_t NP = (size_t)bj_padded(N), nb = NP / kBjB;
    return (3 * NP * NP + (nb / 2) * kBjW * kBjW + 2 * NP + 4) * sizeof(float) + 4 * sizeof(int);
}
// samples processed per pass: bounded by the grid's z extent and by ~6 GB of workspace (the batch is walked in slices)
long long eig_block_slice(long long B, int N) {
    long long s = (long long)(6.0e9 / (double)bj_sample_bytes(N));
    s = s < 64 ? 64 : (s > 16384 ? 16384 : s);
    return B < s ? B : s;
}
size_t eig_block_workspace_bytes(long long B, int N) {
    return (size_t)eig_block_slice(B, N) * bj_sample_bytes(N) + 2048;
}

static BjArgs bj_carve(void *ws, long long B, int N) {
    BjArgs a;
    a.N = N; a.NP = bj_padded(N); a.nb = a.NP / kBjB; a.B = B;
    const size_t NP = (size_t)a.NP, mat = NP * NP;
    unsigned char *p = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    a.state = reinterpret_cast<int *>(p); p += ((size_t)B * 4 * sizeof(int) + 255) / 256 * 256;
    float *f = reinterpret_cast<float *>(p);
    a.S = f; f += (size_t)B * mat;
    a.W = f; f += (size_t)B * mat;
    a.VT = f; f += (size_t)B * mat;
    a.ET = f; f += (size_t)B * (a.nb / 2) * kBjW * kBjW;
    a.mean = f; f += (size_t)B * NP;
    a.rs = f; f += (size_t)B * NP;
    a.trace = f;
    return a;
}

// blocks (bi < bj) of pair `x` in round `round`; round < 0: the neighbour round (2x, 2x + 1)
__device__ __forceinline__ void bj_pair(int nb, int round, int x, int &bi, int &bj) {
    if (round < 0) { bi = 2 * x; bj = 2 * x + 1; return; }
    const int m = nb - 1;
    int a, b;
    if (x == 0) { a = round; b = m; }
    else { a = (round + x) % m; b = (round - x + m) % m; }
    bi = min(a, b); bj = max(a, b);
}

// ---------------------------------------------------------------------------------------------------------------
// column means of the (L, N) observations and, for 'scaled_identity', the mean sample variance (misc.py:157-161)
__global__ void __launch_bounds__(256) cov_stats_kernel(const float *x, int L, int N, int NP, int need_trace, float *mean,
                                                        float *trace) {
    __shared__ float red[8];
    const long long b = blockIdx.x;
    const float *xb = x + b * (long long)L * N;
    float *mb = mean + b * NP;
    float loc = 0.f;
    for (int i = threadIdx.x; i < NP; i += 256) {
        float m = 0.f;
        if (i < N) {
            float a0 = 0.f, a1 = 0.f;
            int l = 0;
            for (; l + 1 < L; l += 2) { a0 += xb[(long long)l * N + i]; a1 += xb[(long long)(l + 1) * N + i]; }
            if (l < L) a0 += xb[(long long)l * N + i];
            m = (a0 + a1) / (float)L;
            if (need_trace) {
                float v = 0.f;
                for (int l2 = 0; l2 < L; ++l2) { const float d = xb[(long long)l2 * N + i] - m; v = fmaf(d, d, v); }
                loc += v;
            }
        }
        mb[i] = m;
    }
    if (need_trace) {
        const float tot = block_sum<float, 256>(loc, red);
        if (threadIdx.x == 0) trace[b] = tot / (float)(L - 1) / (float)N;
    }
}

struct CovEpilogue {      // S_ij = fact * acc * (c | shrinkage weight on the diagonal) + the target's diagonal (misc.py:149-166)
    float *out; int ld, N, lim;
    float woff, wdiag, dadd;
    __device__ __forceinline__ void operator()(int m, int n, float v) const {
        if (m < lim && n < lim) {
            float r = 0.f;
            if (m < N && n < N) r = m == n ? fmaf(v, wdiag, dadd) : v * woff;
            out[(long long)n * ld + m] = r;      // symmetric: (n, m) = (m, n); lanes walk m -> coalesced
        }
    }
};

__global__ void __launch_bounds__(tc::kThreads) cov_gram_tc_kernel(const float *x, const float *coef, int strategy, int L, int N,
                                                                   int NP, const float *mean, const float *trace, float *out,
                                                                   int ld, int lim, long long out_stride) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ tc::TileShared sh;
    const long long b = blockIdx.z;
    const float *xb = x + b * (long long)L * N;
    const float fact = 1.f / (float)(L - 1);
    const float c = strategy == DDW_SHRINK_NONE ? 1.f : coef[b];
    CovEpilogue epi;
    epi.out = out + b * out_stride; epi.ld = ld; epi.N = N; epi.lim = lim;
    epi.woff = fact * c;
    epi.wdiag = strategy == DDW_SHRINK_DIAGONAL ? fact * (c + (1.f - c)) : epi.woff;
    epi.dadd = strategy == DDW_SHRINK_IDENTITY ? 1.f - c : (strategy == DDW_SHRINK_SCALED_IDENTITY ? (1.f - c) * trace[b] : 0.f);
    tc::LoadRowsContig la{xb, N, N, L, mean + b * NP, nullptr};
    tc::tile_gemm<128>(la, la, blockIdx.x * 128, blockIdx.y * 128, L, epi, smem_dyn, sh);
}

// ddw_psd_sqrt_fwd: the input already is the symmetric matrix; S = (m + m') / 2, zero padded
__global__ void __launch_bounds__(256) sym_pad_kernel(const float *m, int N, int NP, float *S) {
    const long long b = blockIdx.y;
    const int i = blockIdx.x;
    const float *mb = m + b * (long long)N * N;
    float *Sb = S + b * (long long)NP * NP + (long long)i * NP;
    for (int j = threadIdx.x; j < NP; j += 256)
        Sb[j] = (i < N && j < N) ? 0.5f * (mb[(long long)i * N + j] + mb[(long long)j * N + i]) : 0.f;
}

__global__ void __launch_bounds__(256) bj_init_kernel(BjArgs a) {
    __shared__ float red[8];
    const long long b = blockIdx.x;
    float *VT = a.VT + b * (long long)a.NP * a.NP;
    const float *S = a.S + b * (long long)a.NP * a.NP;
    float smax = 0.f;
    for (int i = threadIdx.x; i < a.NP; i += 256) {
        VT[(long long)i * a.NP + i] = 1.f;
        smax = fmaxf(smax, fabsf(S[(long long)i * a.NP + i]));
    }
    smax = warp_max(smax);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = smax;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) smax = fmaxf(smax, red[i]);
        a.state[4 * b + 0] = 0; a.state[4 * b + 1] = 0; a.state[4 * b + 2] = 0;
        a.state[4 * b + 3] = __float_as_int(smax);      // scale of the matrix: |trace| / N <= max |diag| for a PSD matrix
    }
}

// start of sweep `sweep` (or, with last != 0, the closing check): a sample whose previous sweep made no significant
// rotation is done
__global__ void __launch_bounds__(256) bj_sweep_begin_kernel(BjArgs a, int sweep, int last) {
    const long long b = (long long)blockIdx.x * 256 + threadIdx.x;
    if (b >= a.B) return;
    int *s = a.state + 4 * b;
    if (s[1]) return;
    if (sweep > 0 && s[0] == 0) { s[1] = 1; return; }
    if (last) return;
    s[0] = 0;
    s[2] = sweep + 1;
}

// ---------------------------------------------------------------------------------------------------------------
// local rotations of one block pair: P = S[IJ, IJ] (64 x 64) in shared memory, U accumulated as rows of UT.
// A step applies 32 disjoint rotations J_k at once.  Indices split into the 32 pairs, so P splits into 32 x 32 blocks of
// 2 x 2 and the two-sided update is  P_kl <- J_k' P_kl J_l : one pass over P per step (lane = l, warps walk k), the lane's own
// (s_l, tau_l) in registers, (s_k, tau_k) broadcast from shared memory -- instead of a column pass and a row pass with a
// barrier between them.  Rotations use Rutishauser's form a' = a - s (b + tau a), b' = b + s (a - tau b).
__global__ void __launch_bounds__(kLocalThreads) bj_local_kernel(BjArgs a, int round) {
    __shared__ float P[kBjW][kBjW + 1];
    __shared__ __align__(8) float UT[kBjW][kBjW + 2];
    __shared__ float cs[kBjB], sn[kBjB];
    __shared__ int pp[kBjB], pq[kBjB];
    const long long b = blockIdx.y;
    if (a.state[4 * b + 1]) return;
    int bi, bj;
    bj_pair(a.nb, round, blockIdx.x, bi, bj);
    const int I0 = bi * kBjB, J0 = bj * kBjB;
    if (I0 >= a.N) return;                 // both blocks are padding: nothing to rotate, the update kernels skip the pair too
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NP = a.NP;
    const float *S = a.S + b * (long long)NP * NP;
    for (int e = tid; e < kBjW * kBjW; e += kLocalThreads) {
        const int i = e >> 6, j = e & 63;
        const int gi = i < kBjB ? I0 + i : J0 + i - kBjB, gj = j < kBjB ? I0 + j : J0 + j - kBjB;
        P[i][j] = S[(long long)gi * NP + gj];
        UT[i][j] = i == j ? 1.f : 0.f;
    }
    __syncthreads();
    const bool full = round < 0;
    const int nsteps = full ? kBjW - 1 : kBjB;
    const float tol = Num<float>::eps();
    const float floor_abs = tol * __int_as_float(a.state[4 * b + 3]);   // eps * max |diag|: below it an entry is rounding noise
    int mysig = 0;
    for (int t = 0; t < nsteps; ++t) {
        int active = 0;
        if (tid < kBjB) {
            const int k = tid;
            int p, q;
            if (full) {
                int x, y;
                if (k == 0) { x = t; y = kBjW - 1; }
                else { x = (t + k) % (kBjW - 1); y = (t - k + kBjW - 1) % (kBjW - 1); }
                p = min(x, y); q = max(x, y);
            } else { p = k; q = kBjB + ((k + t) & (kBjB - 1)); }
            const float spq = P[p][q], spp = P[p][p], sqq = P[q][q];
            const float aq = fabsf(spq), ref = sqrtf(fabsf(spp * sqq));
            float cc = 1.f, ss = 0.f;
            if (aq > tol * ref && aq > floor_abs) {
                const float tau = (sqq - spp) / (2.f * spq);
                const float tt = (tau >= 0.f ? 1.f : -1.f) / (fabsf(tau) + sqrtf(1.f + tau * tau));
                cc = 1.f / sqrtf(1.f + tt * tt);      // correctly rounded (see covariance.cu: rsqrt's bias would shrink V)
                ss = tt * cc;
                active = 1;
                // Jacobi converges quadratically: once every rotated entry of a sweep is below kBjSigRel of its diagonal
                // scale, what is left after the sweep is below rounding and no further sweep is needed
                if (aq > kBjSigRel * ref) ++mysig;
            }
            pp[k] = p; pq[k] = q; cs[k] = ss / (1.f + cc); sn[k] = ss;
        }
        if (!__syncthreads_or(active)) continue;       // no rotation in this step (late sweeps): nothing to apply
        {
            const int pl = pp[lane], ql = pq[lane];
            const float sl = sn[lane], tl = cs[lane];
#pragma unroll 2
            for (int k = warp; k < kBjB; k += kLocalThreads / 32) {
                const int pk = pp[k], qk = pq[k];
                const float sk = sn[k], tk = cs[k];
                float x00 = P[pk][pl], x01 = P[pk][ql], x10 = P[qk][pl], x11 = P[qk][ql];
                // columns (J_l)
                float y00 = fmaf(-sl, fmaf(x00, tl, x01), x00), y01 = fmaf(sl, fmaf(-x01, tl, x00), x01);
                float y10 = fmaf(-sl, fmaf(x10, tl, x11), x10), y11 = fmaf(sl, fmaf(-x11, tl, x10), x11);
                // rows (J_k')
                P[pk][pl] = fmaf(-sk, fmaf(y00, tk, y10), y00); P[qk][pl] = fmaf(sk, fmaf(-y10, tk, y00), y10);
                P[pk][ql] = fmaf(-sk, fmaf(y01, tk, y11), y01); P[qk][ql] = fmaf(sk, fmaf(-y11, tk, y01), y11);
                // rows pk, qk of UT: two columns per lane
                if (sk != 0.f) {
                    float2 u = *reinterpret_cast<float2 *>(&UT[pk][2 * lane]), v = *reinterpret_cast<float2 *>(&UT[qk][2 * lane]);
                    float2 un, vn;
                    un.x = fmaf(-sk, fmaf(u.x, tk, v.x), u.x); vn.x = fmaf(sk, fmaf(-v.x, tk, u.x), v.x);
                    un.y = fmaf(-sk, fmaf(u.y, tk, v.y), u.y); vn.y = fmaf(sk, fmaf(-v.y, tk, u.y), v.y);
                    *reinterpret_cast<float2 *>(&UT[pk][2 * lane]) = un;
                    *reinterpret_cast<float2 *>(&UT[qk][2 * lane]) = vn;
                }
            }
        }
        __syncthreads();
    }
    float *ET = a.ET + (b * (a.nb / 2) + blockIdx.x) * (long long)(kBjW * kBjW);
    for (int e = tid; e < kBjW * kBjW; e += kLocalThreads) {
        const int n = e >> 6, k = e & 63;
        ET[e] = UT[n][k] - (n == k ? 1.f : 0.f);      // B operand of the update GEMMs: (n, k) = U[k][n] - delta
    }
    if (tid < kBjB) {
        mysig = warp_sum(mysig);
        if (tid == 0 && mysig) atomicAdd(&a.state[4 * b], mysig);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// rotation updates on tcgen05: one CTA = 128 columns (m) of the 64 rows IJ:  Y(m, n) = X(m, n) + sum_k X(m, k) E[k][n],
// X(m, k) = Xmat[IJ_k][m] (rows of a row-major matrix: lanes walk m, every load is coalesced).
//   PHASE 0, blockIdx.z even: Xmat = S,  Y -> Z[m][IJ_n]   (Z = (U' S)' : the TRANSPOSE of the row-rotated matrix, 128-byte
//                                                            row pieces per thread)
//   PHASE 0, blockIdx.z odd : Xmat = VT, Y -> VT[IJ_n][m]  in place
//   PHASE 1                 : Xmat = Z,  Y -> S[IJ_n][m]   (S_new = U' S U is symmetric: its rows IJ are U' times the rows IJ
//                                                            of (S U)' ... = U' Z[IJ, :])
// Storing Z transposed makes both phases the same rows-of-a-matrix GEMM.  The CTA is latency bound, so it is kept small
// (55 KB of shared memory, 128 TMEM columns: four CTAs per SM): K = 64 is staged as two chunks of 32 through ONE buffer, the
// loads of the second chunk are in flight while the tensor core works on the first, and the base term X(m, n) is re-read
// (coalesced, from L2) in the epilogue instead of being kept on chip.  Two accumulators (hi*hi and the cross terms): in delta
// form the truncating accumulation acts on the correction only (model: |V'V - I| 4.5e-6 against 3.7e-6 with four).
constexpr int kUpdKC = 32;                                            // k values per staged chunk
constexpr int kUpdSbo = (kUpdKC / 4) * tc::kLbo;                      // 1152
constexpr int kUpdABytes = (tc::kTileM / 8) * kUpdSbo;                // 18432 per hi / lo tile
constexpr int kUpdBBytes = (kBjW / 8) * kUpdSbo;                      // 9216
constexpr size_t kUpdSmem = 2 * kUpdABytes + 2 * kUpdBBytes + 1024;    // 56320: four CTAs per SM

constexpr int kUpdThreads = 128;                                      // one warp per 32 TMEM lanes

template <int PHASE>
__global__ void __launch_bounds__(kUpdThreads, 4) bj_update_kernel(BjArgs a, int round) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ tc::TileShared sh;
    const int which = PHASE == 0 ? (blockIdx.z & 1) : 0;
    const long long b = PHASE == 0 ? (blockIdx.z >> 1) : blockIdx.z;
    if (a.state[4 * b + 1]) return;
    int bi, bj;
    bj_pair(a.nb, round, blockIdx.y, bi, bj);
    const int I0 = bi * kBjB, J0 = bj * kBjB;
    if (I0 >= a.N) return;
    const int NP = a.NP, m0 = blockIdx.x * tc::kTileM;
    const int tid = threadIdx.x, warp = tid >> 5;
    const long long mat = (long long)NP * NP;
    const float *X = (PHASE == 0 ? (which ? a.VT : a.S) : a.W) + b * mat;
    float *Y = (PHASE == 0 ? (which ? a.VT : a.W) : a.S) + b * mat;
    const float *ET = a.ET + (b * (a.nb / 2) + blockIdx.y) * (long long)(kBjW * kBjW);
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *a_hi = base, *a_lo = base + kUpdABytes, *b_hi = base + 2 * kUpdABytes, *b_lo = b_hi + kUpdBBytes;
    constexpr uint32_t kCols = 2 * kBjW;      // 128: accumulator 0 = hi*hi, accumulator 1 = cross terms
    constexpr uint32_t idesc = tc::idesc_tf32(tc::kTileM, kBjW);

    // element (m, k): row IJ_k of X, column m0 + m.  Thread -> m = tid (its TMEM lane as well) and ALL 64 k: the values it
    // stages are also the base term X(m, n) of its own epilogue row, so they stay in registers and nothing is re-read.
    float va[16][4], vb[4][4];
#pragma unroll
    for (int it = 0; it < 16; ++it) {         // 64 independent 4-byte loads in flight per thread
        const int k = 4 * it;
        const int r0 = k < kBjB ? I0 + k : J0 + k - kBjB;
        const float *src = X + (long long)r0 * NP + m0 + tid;
#pragma unroll
        for (int q = 0; q < 4; ++q) va[it][q] = __ldg(src + (long long)q * NP);
    }
    auto load_e = [&](int c) {               // E^T: (n, k) = ET[n][k]; thread -> (k4 = id % 8, n = id / 8)
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int id = it * kUpdThreads + tid;
            const float4 t4 = __ldg(reinterpret_cast<const float4 *>(ET + (id >> 3) * kBjW + c * kUpdKC + 4 * (id & 7)));
            vb[it][0] = t4.x; vb[it][1] = t4.y; vb[it][2] = t4.z; vb[it][3] = t4.w;
        }
    };
    auto store_chunk = [&](int c) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            float4 h, l;
            tc::split_tf32(va[8 * c + it][0], h.x, l.x); tc::split_tf32(va[8 * c + it][1], h.y, l.y);
            tc::split_tf32(va[8 * c + it][2], h.z, l.z); tc::split_tf32(va[8 * c + it][3], h.w, l.w);
            const int off = (tid >> 3) * kUpdSbo + it * tc::kLbo + (tid & 7) * 16;
            *reinterpret_cast<float4 *>(a_hi + off) = h;
            *reinterpret_cast<float4 *>(a_lo + off) = l;
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int id = it * kUpdThreads + tid, n = id >> 3, k4 = id & 7;
            float4 h, l;
            tc::split_tf32(vb[it][0], h.x, l.x); tc::split_tf32(vb[it][1], h.y, l.y);
            tc::split_tf32(vb[it][2], h.z, l.z); tc::split_tf32(vb[it][3], h.w, l.w);
            const int off = (n >> 3) * kUpdSbo + k4 * tc::kLbo + (n & 7) * 16;
            *reinterpret_cast<float4 *>(b_hi + off) = h;
            *reinterpret_cast<float4 *>(b_lo + off) = l;
        }
    };
    auto issue_chunk = [&](int c, uint32_t tmem) {
        const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
        for (int kk = 0; kk < kUpdKC / 8; ++kk) {
            const uint32_t o = kk * 2 * tc::kLbo;
            const uint64_t dah = tc::smem_desc(ah + o, tc::kLbo, kUpdSbo), dal = tc::smem_desc(al + o, tc::kLbo, kUpdSbo);
            const uint64_t dbh = tc::smem_desc(bh + o, tc::kLbo, kUpdSbo), dbl = tc::smem_desc(bl + o, tc::kLbo, kUpdSbo);
            tc::mma_tf32(tmem + kBjW, dal, dbh, idesc, (c | kk) != 0);
            tc::mma_tf32(tmem + kBjW, dah, dbl, idesc, 1u);
            tc::mma_tf32(tmem, dah, dbh, idesc, (c | kk) != 0);
        }
    };

    load_e(0);
    if (tid == 0) { mbar_init(&sh.done[0], 1); mbar_init(&sh.done[1], 1); fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&sh.tmem_base, kCols);
    store_chunk(0);
    load_e(1);                                // in flight while the tensor core works on chunk 0
    fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = sh.tmem_base;
    if (tid == 0) { issue_chunk(0, tmem); tc::mma_commit(&sh.done[0]); }
    mbar_wait(&sh.done[0], 0);                // chunk 0 consumed: the buffer is free again
    store_chunk(1);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) { tc::fence_after_sync(); issue_chunk(1, tmem); tc::mma_commit(&sh.done[1]); }
    mbar_wait(&sh.done[1], 0);
    tc::fence_after_sync();
    // epilogue: this thread owns row m = tid; four quarters of 16 columns keep the register count at 128
#pragma unroll
    for (int qt = 0; qt < 4; ++qt) {
        const int r0 = (qt < 2 ? I0 : J0) + 16 * (qt & 1);
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(16 * qt);
        float v[16], x[16];
        tc::tmem_ld16(lane_addr, v);
        tc::tmem_ld16(lane_addr + kBjW, x);
        if (PHASE == 0 && which == 0) {
            // Z[m][IJ_n]: 16 consecutive floats of row m
            float4 *dst = reinterpret_cast<float4 *>(Y + (long long)(m0 + tid) * NP + r0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                dst[j] = make_float4(va[4 * qt + j][0] + (v[4 * j] + x[4 * j]), va[4 * qt + j][1] + (v[4 * j + 1] + x[4 * j + 1]),
                                     va[4 * qt + j][2] + (v[4 * j + 2] + x[4 * j + 2]), va[4 * qt + j][3] + (v[4 * j + 3] + x[4 * j + 3]));
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) Y[(long long)(r0 + j) * NP + m0 + tid] = va[4 * qt + (j >> 2)][j & 3] + (v[j] + x[j]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, kCols);
}

// ---------------------------------------------------------------------------------------------------------------
// eigenvalues -> sqrt|lambda| with the reference's rank threshold s > s.max * N * eps (misc.py:185-199)
__global__ void __launch_bounds__(256) bj_eigs_kernel(BjArgs a, float *eigval, int32_t *sweeps) {
    __shared__ float red[8];
    const long long b = blockIdx.x;
    const int NP = a.NP, N = a.N, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *S = a.S + b * (long long)NP * NP;
    float smax = 0.f;
    for (int i = threadIdx.x; i < N; i += 256) smax = fmaxf(smax, fabsf(S[(long long)i * NP + i]));
    smax = warp_max(smax);
    if (lane == 0) red[warp] = smax;
    __syncthreads();
    smax = 0.f;
    for (int i = 0; i < 8; ++i) smax = fmaxf(smax, red[i]);
    const float thr = smax * (float)N * Num<float>::eps();
    for (int i = threadIdx.x; i < NP; i += 256) {
        float r = 0.f;
        if (i < N) { const float sv = fabsf(S[(long long)i * NP + i]); r = sv > thr ? sqrtf(sv) : 0.f; }
        a.rs[b * NP + i] = r;
        if (eigval && i < N) eigval[b * N + i] = r;
    }
    if (sweeps && threadIdx.x == 0) sweeps[b] = a.state[4 * b + 1] ? a.state[4 * b + 2] : -a.state[4 * b + 2];   // < 0: not converged
}

__global__ void __launch_bounds__(256) bj_copy_vt_kernel(BjArgs a, float *eigvec) {
    const long long b = blockIdx.y;
    const int i = blockIdx.x;
    const float *src = a.VT + b * (long long)a.NP * a.NP + (long long)i * a.NP;
    float *dst = eigvec + b * (long long)a.N * a.N + (long long)i * a.N;
    for (int j = threadIdx.x; j < a.N; j += 256) dst[j] = src[j];
}

struct StoreSym {
    float *out; int N;
    __device__ __forceinline__ void operator()(int m, int n, float v) const { if (m < N && n < N) out[(long long)n * N + m] = v; }
};

// out = sum_k r_k v_k v_k'  with v_k = row k of VT
__global__ void __launch_bounds__(tc::kThreads) recon_tc_kernel(BjArgs a, float *out) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ tc::TileShared sh;
    const long long b = blockIdx.z;
    const float *VT = a.VT + b * (long long)a.NP * a.NP;
    tc::LoadRowsContig la{VT, a.NP, a.N, a.N, nullptr, a.rs + b * a.NP};
    tc::LoadRowsContig lb{VT, a.NP, a.N, a.N, nullptr, nullptr};
    tc::tile_gemm<128>(la, lb, blockIdx.x * 128, blockIdx.y * 128, a.N, StoreSym{out + b * (long long)a.N * a.N, a.N}, smem_dyn, sh);
}

// ---------------------------------------------------------------------------------------------------------------
static int configure_tc_kernels() {
    static DeviceOnce configured;
    const int dev = current_device();
    if (!configured.needs(dev)) return 0;
    if (cudaFuncSetAttribute(cov_gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::tile_gemm_smem<128>()) != cudaSuccess ||
        cudaFuncSetAttribute(recon_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::tile_gemm_smem<128>()) != cudaSuccess ||
        cudaFuncSetAttribute(bj_update_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpdSmem) != cudaSuccess ||
        cudaFuncSetAttribute(bj_update_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpdSmem) != cudaSuccess)
        return check_launch("cudaFuncSetAttribute(tensor-core covariance kernels)");
    configured.mark(dev);
    return 0;
}

// float32 CovarianceMatrix forward for n_assets beyond shared memory.  L == 0: x is the (B, N, N) matrix itself (ddw_psd_sqrt_fwd).
static int covariance_fwd_large_slice(const float *x, const float *coef, int strategy, int do_sqrt, long long B, int L, int N,
                                      float *out, float *eigvec, float *eigval, int32_t *sweeps, void *ws, cudaStream_t st);

int covariance_fwd_large_f32(const float *x, const float *coef, int strategy, int do_sqrt, long long B, int L, int N, float *out,
                             float *eigvec, float *eigval, int32_t *sweeps, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (!ws || ws_bytes < eig_block_workspace_bytes(B, N)) { set_error("ddw_covariance_fwd: workspace too small (%zu < %zu)", ws_bytes, eig_block_workspace_bytes(B, N)); return DDW_ERR_WORKSPACE; }
    int rc = configure_tc_kernels();
    if (rc) return rc;
    const long long slice = eig_block_slice(B, N), NN = (long long)N * N, xin = L > 0 ? (long long)L * N : NN;
    for (long long b0 = 0; b0 < B; b0 += slice) {
        const long long nb = B - b0 < slice ? B - b0 : slice;
        rc = covariance_fwd_large_slice(x + b0 * xin, coef ? coef + b0 : nullptr, strategy, do_sqrt, nb, L, N, out + b0 * NN,
                                        eigvec ? eigvec + b0 * NN : nullptr, eigval ? eigval + b0 * N : nullptr,
                                        sweeps ? sweeps + b0 : nullptr, ws, st);
        if (rc) return rc;
    }
    return 0;
}

static int covariance_fwd_large_slice(const float *x, const float *coef, int strategy, int do_sqrt, long long B, int L, int N,
                                      float *out, float *eigvec, float *eigval, int32_t *sweeps, void *ws, cudaStream_t st) {
    int rc;
    BjArgs a = bj_carve(ws, B, N);
    const int NP = a.NP;
    const long long mat = (long long)NP * NP;
    const size_t gsm = tc::tile_gemm_smem<128>();
    if (L > 0) {
        cov_stats_kernel<<<(unsigned)B, 256, 0, st>>>(x, L, N, NP, strategy == DDW_SHRINK_SCALED_IDENTITY, a.mean, a.trace);
        if (!do_sqrt) {
            const unsigned t = (unsigned)((N + 127) / 128);
            cov_gram_tc_kernel<<<dim3(t, t, (unsigned)B), tc::kThreads, gsm, st>>>(x, coef, strategy, L, N, NP, a.mean, a.trace, out, N, N, (long long)N * N);
            return check_launch("cov_gram_tc_kernel");
        }
        const unsigned t = (unsigned)(NP / 128);
        cov_gram_tc_kernel<<<dim3(t, t, (unsigned)B), tc::kThreads, gsm, st>>>(x, coef, strategy, L, N, NP, a.mean, a.trace, a.S, NP, NP, mat);
    } else {
        sym_pad_kernel<<<dim3((unsigned)NP, (unsigned)B), 256, 0, st>>>(x, N, NP, a.S);
    }
    if ((rc = check_launch("covariance (large) Gram"))) return rc;
    cudaMemsetAsync(a.W, 0, sizeof(float) * (size_t)B * mat, st);
    cudaMemsetAsync(a.VT, 0, sizeof(float) * (size_t)B * mat, st);
    bj_init_kernel<<<(unsigned)B, 256, 0, st>>>(a);
    const unsigned chunks = (unsigned)(NP / 128), pairs = (unsigned)(a.nb / 2);
    for (int sweep = 0; sweep < kBjMaxSweeps; ++sweep) {
        bj_sweep_begin_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(a, sweep, 0);
        for (int round = -1; round < a.nb - 1; ++round) {
            bj_local_kernel<<<dim3(pairs, (unsigned)B), kLocalThreads, 0, st>>>(a, round);
            bj_update_kernel<0><<<dim3(chunks, pairs, (unsigned)(2 * B)), kUpdThreads, kUpdSmem, st>>>(a, round);
            bj_update_kernel<1><<<dim3(chunks, pairs, (unsigned)B), kUpdThreads, kUpdSmem, st>>>(a, round);
        }
        if ((rc = check_launch("blocked Jacobi sweep"))) return rc;
    }
    bj_sweep_begin_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(a, kBjMaxSweeps, 1);   // marks samples whose last sweep was clean
    bj_eigs_kernel<<<(unsigned)B, 256, 0, st>>>(a, eigval, sweeps);
    if (eigvec) bj_copy_vt_kernel<<<dim3((unsigned)N, (unsigned)B), 256, 0, st>>>(a, eigvec);
    const unsigned t = (unsigned)((N + 127) / 128);
    recon_tc_kernel<<<dim3(t, t, (unsigned)B), tc::kThreads, gsm, st>>>(a, out);
    return check_launch("recon_tc_kernel");
}

}  // namespace ddw
