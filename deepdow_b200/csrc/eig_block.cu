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
//   bj_update_kernel<0>  W[IJ, :] = U' S[IJ, :] and VT[IJ, :] = U' VT[IJ, :]   (128 x 64 x 64 tiles on tcgen05)
//   bj_update_kernel<1>  S[IJ, :] = U' W'[IJ, :]                                (S stays symmetric: rows IJ of S' = columns IJ of S)
// The updates run in "delta" form X + X E: the tensor core only forms the correction, which vanishes as the sweeps converge,
// and the base term is added back exactly in FP32 (numerical model: tools/models/block_jacobi_model.py).
// All samples of the batch advance together (grid = pairs x batch); a sample whose last sweep made no significant rotation is
// flagged done and its CTAs return at once.  Matrices live in the caller's workspace, padded to a multiple of 128.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace ddw {

constexpr int kBjB = 32;              // block width
constexpr int kBjW = 2 * kBjB;        // width of a block pair = K and N of the update GEMMs
constexpr int kBjMaxSweeps = 24;
constexpr int kLocalThreads = 128;

struct BjArgs {
    int N, NP, nb;                    // n_assets, padded size (multiple of 128), number of 32-wide blocks (NP / 32)
    long long B;
    float *S, *W, *VT, *ET;           // (B, NP, NP) x 3, (B, nb/2, 64, 64)
    float *mean, *rs, *trace;         // (B, NP), (B, NP), (B)
    int *state;                       // (B, 4): [0] significant rotations of the running sweep, [1] done, [2] sweeps used
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
    const long long b = blockIdx.x;
    float *VT = a.VT + b * (long long)a.NP * a.NP;
    for (int i = threadIdx.x; i < a.NP; i += 256) VT[(long long)i * a.NP + i] = 1.f;
    if (threadIdx.x < 4) a.state[4 * b + threadIdx.x] = 0;
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
// local rotations of one block pair: P = S[IJ, IJ] (64 x 64) in shared memory, U accumulated as rows of UT
__global__ void __launch_bounds__(kLocalThreads) bj_local_kernel(BjArgs a, int round) {
    __shared__ float P[kBjW][kBjW + 1];
    __shared__ float UT[kBjW][kBjW + 1];
    __shared__ float cs[kBjB], sn[kBjB];
    __shared__ int pp[kBjB], pq[kBjB];
    __shared__ int nsig;
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
    if (tid == 0) nsig = 0;
    __syncthreads();
    const bool full = round < 0;
    const int nsteps = full ? kBjW - 1 : kBjB;
    const float tol = Num<float>::eps();
    int mysig = 0;
    for (int t = 0; t < nsteps; ++t) {
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
            if (aq > tol * ref && aq > Num<float>::tiny()) {
                const float tau = (sqq - spp) / (2.f * spq);
                const float tt = (tau >= 0.f ? 1.f : -1.f) / (fabsf(tau) + sqrtf(1.f + tau * tau));
                cc = 1.f / sqrtf(1.f + tt * tt);      // correctly rounded (see covariance.cu: rsqrt's bias would shrink V)
                ss = tt * cc;
                if (aq > 8.f * tol * ref) ++mysig;
            }
            pp[k] = p; pq[k] = q; cs[k] = ss / (1.f + cc); sn[k] = ss;
        }
        __syncthreads();
        // columns p, q of P and rows p, q of UT: Rutishauser's form a' = a - s (b + tau a), b' = b + s (a - tau b)
        for (int k = warp; k < kBjB; k += kLocalThreads / 32) {
            const float ss = sn[k];
            if (ss == 0.f) continue;
            const float tt = cs[k];
            const int p = pp[k], q = pq[k];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = lane + 32 * h;
                const float x = P[i][p], y = P[i][q];
                P[i][p] = fmaf(-ss, fmaf(x, tt, y), x);
                P[i][q] = fmaf(ss, fmaf(-y, tt, x), y);
                const float ux = UT[p][i], uy = UT[q][i];
                UT[p][i] = fmaf(-ss, fmaf(ux, tt, uy), ux);
                UT[q][i] = fmaf(ss, fmaf(-uy, tt, ux), uy);
            }
        }
        __syncthreads();
        for (int k = warp; k < kBjB; k += kLocalThreads / 32) {
            const float ss = sn[k];
            if (ss == 0.f) continue;
            const float tt = cs[k];
            const int p = pp[k], q = pq[k];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = lane + 32 * h;
                const float x = P[p][j], y = P[q][j];
                P[p][j] = fmaf(-ss, fmaf(x, tt, y), x);
                P[q][j] = fmaf(ss, fmaf(-y, tt, x), y);
            }
        }
        __syncthreads();
    }
    if (mysig) atomicAdd(&nsig, mysig);
    float *ET = a.ET + (b * (a.nb / 2) + blockIdx.x) * (long long)(kBjW * kBjW);
    for (int e = tid; e < kBjW * kBjW; e += kLocalThreads) {
        const int n = e >> 6, k = e & 63;
        ET[e] = UT[n][k] - (n == k ? 1.f : 0.f);      // B operand of the update GEMMs: (n, k) = U[k][n] - delta
    }
    __syncthreads();
    if (tid == 0 && nsig) atomicAdd(&a.state[4 * b], nsig);
}

// ---------------------------------------------------------------------------------------------------------------
// rotation updates on tcgen05: one CTA = 128 columns (m) of the 64 rows IJ; out[IJ_n][m] = X(m, n) + sum_k X(m, k) E[k][n]
//   PHASE 0: X(m, k) = S[IJ_k][m] -> W  (blockIdx.z even)   and   X(m, k) = VT[IJ_k][m] -> VT in place  (blockIdx.z odd)
//   PHASE 1: X(m, k) = W[m][IJ_k] -> S
constexpr int kUpdSbo = (kBjW / 4) * tc::kLbo;                       // 2304
constexpr int kUpdABytes = (tc::kTileM / 8) * kUpdSbo;               // 36864 per hi / lo tile
constexpr int kUpdBBytes = (kBjW / 8) * kUpdSbo;                     // 18432
constexpr size_t kUpdSmem = 2 * kUpdABytes + 2 * kUpdBBytes + 1024;   // ~109 KB: two CTAs per SM

struct LoadPairRows {        // element (m, k) = X[row(k)][m]; lanes walk m
    static constexpr bool kRowsFast = true;
    const float *X; int NP, I0, J0;
    __device__ __forceinline__ void ld4(int m, int k, float (&v)[4]) const {
        const int r0 = k < kBjB ? I0 + k : J0 + k - kBjB;
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = X[(long long)(r0 + q) * NP + m];
    }
};
struct LoadPairCols {        // element (m, k) = X[m][col(k)]; lanes walk k
    static constexpr bool kRowsFast = false;
    const float *X; int NP, I0, J0;
    __device__ __forceinline__ void ld4(int m, int k, float (&v)[4]) const {
        const int c0 = k < kBjB ? I0 + k : J0 + k - kBjB;
        const float4 t = *reinterpret_cast<const float4 *>(X + (long long)m * NP + c0);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};
struct LoadE {               // element (n, k) = ET[n][k]
    static constexpr bool kRowsFast = false;
    const float *ET;
    __device__ __forceinline__ void ld4(int n, int k, float (&v)[4]) const {
        const float4 t = *reinterpret_cast<const float4 *>(ET + n * kBjW + k);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};

template <int PHASE>
__global__ void __launch_bounds__(tc::kThreads) bj_update_kernel(BjArgs a, int round) {
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
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long mat = (long long)NP * NP;
    const float *X = PHASE == 0 ? (which ? a.VT : a.S) + b * mat : a.W + b * mat;
    float *Y = PHASE == 0 ? (which ? a.VT : a.W) + b * mat : a.S + b * mat;
    const float *ET = a.ET + (b * (a.nb / 2) + blockIdx.y) * (long long)(kBjW * kBjW);
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *a_hi = base, *a_lo = base + kUpdABytes, *b_hi = base + 2 * kUpdABytes, *b_lo = b_hi + kUpdBBytes;
    constexpr uint32_t kCols = tc::kAccs * kBjW;      // 256
    if (tid == 0) { mbar_init(&sh.fin, 1); fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&sh.tmem_base, kCols);
    if (PHASE == 0) tc::stage_operand<tc::kTileM, kBjW>(LoadPairRows{X, NP, I0, J0}, m0, 0, a_hi, a_lo);
    else tc::stage_operand<tc::kTileM, kBjW>(LoadPairCols{X, NP, I0, J0}, m0, 0, a_hi, a_lo);
    tc::stage_operand<kBjW, kBjW>(LoadE{ET}, 0, 0, b_hi, b_lo);
    fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = sh.tmem_base;
    if (tid == 0) {
        tc::issue_mma3<kBjW>(tmem, smem_u32(a_hi), smem_u32(a_lo), smem_u32(b_hi), smem_u32(b_lo), kBjW / 8, 0, tc::kLbo, kUpdSbo);
        tc::mma_commit(&sh.fin);
    }
    mbar_wait(&sh.fin, 0);
    tc::fence_after_sync();
    {
        const int ml = 32 * (warp & 3) + lane, c0 = 32 * (warp >> 2);
        float v[32];
        tc::load_acc32<kBjW>(tmem, warp, c0, v);
        const int aoff = (ml >> 3) * kUpdSbo + (ml & 7) * 16;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int n = c0 + j;
            const int off = aoff + (n >> 2) * tc::kLbo + (n & 3) * 4;
            const float basev = *reinterpret_cast<const float *>(a_hi + off) + *reinterpret_cast<const float *>(a_lo + off);
            const int row = n < kBjB ? I0 + n : J0 + n - kBjB;
            Y[(long long)row * NP + m0 + ml] = basev + v[j];
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
            bj_update_kernel<0><<<dim3(chunks, pairs, (unsigned)(2 * B)), tc::kThreads, kUpdSmem, st>>>(a, round);
            bj_update_kernel<1><<<dim3(chunks, pairs, (unsigned)B), tc::kThreads, kUpdSmem, st>>>(a, round);
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
