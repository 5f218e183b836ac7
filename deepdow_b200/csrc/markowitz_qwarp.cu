// markowitz_qwarp.cu -- float32 NumericalMarkowitz forward for n_assets <= 128 as two kernels:
//
//   (1) markowitz_gram_small_kernel   Q = A'A of every portfolio on the 5th-generation tensor cores
//       (tcgen05.mma.kind::tf32 in the 3xTF32 split, accumulators in TMEM).  A = gamma-tiled covmat_sqrt
//       (reference deepdow/layers/allocate.py:268-273) arrives by ONE bulk TMA copy per portfolio (4 N^2 bytes,
//       double buffered), is split into TF32 hi/lo parts chunk by chunk into the tensor core's canonical K-major
//       layout, and Q leaves through tcgen05.ld as coalesced stores into a scratch that stays in the 126 MB L2.
//   (2) markowitz_qwarp_kernel        the primal-dual active-set iteration of markowitz_sparse.cuh with ONE WARP per
//       portfolio and no shared copy of A or Q: the reduced KKT system of the free assets (<= 32) lives in registers,
//       one row per lane, and is eliminated with warp shuffles; gradients read the few rows of Q they need from L2.
//
// Why: the one-CTA-per-portfolio kernel keeps 40 KB of A resident per portfolio (4 portfolios per SM) while ~60% of its
// time is spent in phases that never touch A, and spends ~37 k warp instructions per portfolio, most of them index
// arithmetic and barriers around ~2.5 k warp instructions of useful FMAs.  With Q formed by the tensor core the solver
// needs only |F| rows of Q per iteration; one warp per portfolio removes every block barrier and lets 16 portfolios
// share an SM.  Portfolios whose free set outgrows 32 assets are handed to the dense kernel (work_list), those that do
// not converge in kMaxPdas iterations to the interior-point kernel (fail_list) -- the same protocol as the
// one-CTA-per-portfolio kernel, which stays the float64 path.
#include <stdlib.h>
#include "markowitz_common.cuh"
#include "support_kernels.cuh"
#include "tc_gemm.cuh"

namespace ddw {

// ------------------------------------------------------------------------------------------------------------------
// (1) Gram kernel
// ------------------------------------------------------------------------------------------------------------------
constexpr int kGsThreads = 256;
constexpr int kGsChunkK = 16;                       // k values per staged chunk = 2 MMA k-steps
constexpr int kGsLbo = 128;                         // K-adjacent core matrices (8 rows x 16 B) are contiguous
constexpr int kGsSbo = (kGsChunkK / 4) * kGsLbo + 16;   // 528 B between 8-row groups: the 16 B skew makes the staging stores conflict-free

struct GsLayout {
    int BN;                 // MMA N extent = n_assets rounded up to 16
    int nraw;               // raw A buffers (2 when they fit next to a second CTA on the SM)
    unsigned raw_bytes;     // one raw buffer, rounded to 128 B
    unsigned tile_bytes;    // one hi or lo tile: BN rows x 16 k
    unsigned total;         // dynamic shared memory of the kernel
    unsigned tmem_cols;
};
static inline GsLayout gs_layout(int N, int smem_per_sm) {
    GsLayout l;
    l.BN = (N + 15) / 16 * 16;
    l.raw_bytes = (unsigned)(((size_t)N * N * 4 + 127) / 128 * 128);
    l.tile_bytes = (unsigned)(l.BN / 8) * kGsSbo;
    // the M = 128 operand reads 16 row groups of every tile: (16 - BN/8) groups past its end.  They land in the next tile of
    // the ring (finite or not, they only reach output rows >= BN, which nobody reads); the pad keeps the last tile's
    // over-read inside the allocation
    const unsigned ring = 4 * l.tile_bytes + (unsigned)(16 - l.BN / 8) * kGsSbo;
    const unsigned fixed = 1024 /* alignment slack */ + 128 /* barriers */ + 2048 /* partial column norms */ + ring;
    l.nraw = (2 * (fixed + 2 * l.raw_bytes + 1024) <= (unsigned)smem_per_sm) ? 2 : 1;
    l.total = fixed + l.nraw * l.raw_bytes;
    const unsigned need = 2u * (unsigned)l.BN;
    l.tmem_cols = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : 256));
    return l;
}

struct GsParams {
    const float *C, *gamma, *alpha;
    float *Q, *diag;
    long long B, Bmod, b0;
    int N, LDQ, BN, nraw;
    unsigned raw_bytes, tile_bytes, tmem_cols;
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}

// a == hi + lo exactly, hi = a rounded to TF32 by integer arithmetic (two instructions; cvt.rna.tf32 is emulated with four
// on this target): add half an ulp of the 10-bit mantissa, clear the low 13 bits.  Non-finite inputs stay non-finite.
__device__ __forceinline__ void split_fast(float a, float &hi, float &lo) {
    hi = __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u);
    lo = a - hi;
}

// gamma tiling (allocate.py:268-270) in place, As[e] *= gamma[(gflat0 + e) % Bmod], with the gamma loads of four
// 16-byte groups in flight at once (apply_gamma_tiling waits for every load before it issues the next)
__device__ __forceinline__ void gamma_tiling_batched(float *As, int N, const float *__restrict__ gamma, long long gflat0, long long Bmod) {
    const int tid = threadIdx.x;
    const unsigned uB = (unsigned)Bmod;
    if ((N & 3) == 0 && (uB & 3u) == 0u) {
        unsigned gi = (unsigned)((gflat0 + 4 * tid) % Bmod);
        const unsigned gstep = (unsigned)((4 * kGsThreads) % Bmod);
        const int NN = N * N;
        for (int e0 = 4 * tid; e0 < NN; e0 += 16 * kGsThreads) {
            float4 g[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (e0 + q * 4 * kGsThreads < NN) g[q] = __ldg(reinterpret_cast<const float4 *>(gamma + gi));
                gi += gstep; if (gi >= uB) gi -= uB;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int e = e0 + q * 4 * kGsThreads;
                if (e < NN) {
                    float4 a = *reinterpret_cast<float4 *>(As + e);
                    a.x *= g[q].x; a.y *= g[q].y; a.z *= g[q].z; a.w *= g[q].w;
                    *reinterpret_cast<float4 *>(As + e) = a;
                }
            }
        }
    } else {
        unsigned gi = (unsigned)((gflat0 + tid) % Bmod);
        const unsigned gstep = (unsigned)(kGsThreads % Bmod);
        for (int e = tid; e < N * N; e += kGsThreads) {
            As[e] *= __ldg(gamma + gi);
            gi += gstep; if (gi >= uB) gi -= uB;
        }
    }
}

__global__ void __launch_bounds__(kGsThreads, 2) markowitz_gram_small_kernel(GsParams p) {
    extern __shared__ unsigned char gs_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, BN = p.BN;
    // 1 KB alignment by an offset from the shared array itself: the pointers keep their address space (LDS / STS, not generic)
    unsigned char *base = gs_smem + ((1024u - (smem_u32(gs_smem) & 1023u)) & 1023u);
    unsigned long long *bar_raw = reinterpret_cast<unsigned long long *>(base);          // [2]
    unsigned long long *bar_done = bar_raw + 2;                                          // [2] ring stage drained
    unsigned long long *bar_fin = bar_raw + 4;                                           // accumulators complete
    uint32_t &tmem_slot = *reinterpret_cast<uint32_t *>(base + 64);
    float *dpart = reinterpret_cast<float *>(base + 128);                                // [4][128] partial column norms
    unsigned char *ring = base + 128 + 2048;
    const unsigned stage_bytes = 2 * p.tile_bytes;
    unsigned char *raw0 = ring + 2 * stage_bytes + (unsigned)(16 - BN / 8) * kGsSbo;
    const unsigned nn_bytes = (unsigned)(N * N * 4);
    const bool use_tma = (nn_bytes & 15u) == 0u && (reinterpret_cast<unsigned long long>(p.C) & 15ull) == 0ull;

    if (tid == 0) {
        mbar_init(&bar_raw[0], 1); mbar_init(&bar_raw[1], 1); mbar_init(&bar_done[0], 1); mbar_init(&bar_done[1], 1);
        mbar_init(bar_fin, 1);
        fence_mbar_init();
    }
    // rows >= n_assets of the tiles are never written again: zero the whole ring once
    for (unsigned e = tid * 16u; e < 2 * stage_bytes + (unsigned)(16 - BN / 8) * kGsSbo; e += kGsThreads * 16u)
        *reinterpret_cast<float4 *>(ring + e) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == 0) tc::tmem_alloc(&tmem_slot, p.tmem_cols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = tc::idesc_tf32(128, BN);

    const long long b_first = blockIdx.x, b_step = gridDim.x;
    // prefetch the first (two) portfolios
    if (use_tma && tid == 0) {
        for (int q = 0; q < p.nraw; ++q) {
            const long long b = b_first + q * b_step;
            if (b < p.B) {
                mbar_expect_tx(&bar_raw[q], nn_bytes);
                tma_bulk_g2s(raw0 + (size_t)q * p.raw_bytes, p.C + b * (long long)N * N, nn_bytes, &bar_raw[q]);
            }
        }
    }
    const int nchunks = (N + kGsChunkK - 1) / kGsChunkK;
    unsigned gchunk = 0;          // chunks staged so far by this CTA (ring position)
    long long it = 0;             // portfolios processed so far by this CTA
    // staging.  Fast path (n_assets % 4 == 0): thread = (quad of operand rows 4 rq .. 4 rq + 3, k group k4 of the chunk): four
    // 16-byte loads of A (one per k), split, four 16-byte stores per tile (one per row) -- the 4 x 4 transpose is register
    // naming only.  Generic path: thread = (row r, k groups kh and kh + 2), scalar loads.
    const bool quads = (N & 3) == 0;
    const int N4 = N >> 2;
    const int rq = quads ? tid % max(N4, 1) : 0, qk4 = quads ? tid / max(N4, 1) : 4;
    const int r = tid & 127, kh = tid >> 7;
    const unsigned soff = (unsigned)((r >> 3) * kGsSbo + (r & 7) * 16 + kh * kGsLbo);
    const unsigned qoff = (unsigned)((rq >> 1) * kGsSbo + (rq & 1) * 64 + qk4 * kGsLbo);
    // epilogue: this thread's TMEM lane = row m of Q, the two warp groups split the columns
    const int m = 32 * (warp & 3) + lane;
    const int half = (BN / 2 + 7) / 8 * 8;
    const int cbeg = (warp >> 2) * half, cend = min(min(BN, cbeg + half), (N + 7) / 8 * 8);
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    for (long long b = b_first; b < p.B; b += b_step, ++it) {
        const int rb = p.nraw == 2 ? (int)(it & 1) : 0;
        float *raw = reinterpret_cast<float *>(raw0 + (size_t)rb * p.raw_bytes);
        if (use_tma) mbar_wait(&bar_raw[rb], (unsigned)((p.nraw == 2 ? it >> 1 : it) & 1));
        else {
            const float *Cb = p.C + b * (long long)N * N;
            for (int e = tid; e < N * N; e += kGsThreads) raw[e] = Cb[e];
            __syncthreads();
        }
        gamma_tiling_batched(raw, N, p.gamma, (p.b0 + b) * (long long)N * N, p.Bmod);
        __syncthreads();
        float dacc[4] = {0.f, 0.f, 0.f, 0.f};        // sum of squares of this thread's elements: partial diag(Q) in plain FP32
        for (int c = 0; c < nchunks; ++c, ++gchunk) {
            const unsigned s = gchunk & 1u;
            unsigned char *hi = ring + s * stage_bytes, *lo = hi + p.tile_bytes;
            const int kleft = N - c * kGsChunkK;                 // k values of this chunk that exist
            const int ksteps = kleft > 8 ? 2 : 1;                // 8-wide MMA k-steps to issue
            if (gchunk >= 2) mbar_wait(&bar_done[s], ((gchunk >> 1) + 1u) & 1u);     // the MMAs that read this stage have drained
            if (quads) {
                if (qk4 < 2 * ksteps) {
                    float4 a[4];
                    const int k = c * kGsChunkK + 4 * qk4;
                    if (k < N) {
                        const float *src = raw + k * N + 4 * rq;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) a[kk] = *reinterpret_cast<const float4 *>(src + kk * N);
                    } else {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) a[kk] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 vh, vl;
                        const float e0 = q == 0 ? a[0].x : (q == 1 ? a[0].y : (q == 2 ? a[0].z : a[0].w));
                        const float e1 = q == 0 ? a[1].x : (q == 1 ? a[1].y : (q == 2 ? a[1].z : a[1].w));
                        const float e2 = q == 0 ? a[2].x : (q == 1 ? a[2].y : (q == 2 ? a[2].z : a[2].w));
                        const float e3 = q == 0 ? a[3].x : (q == 1 ? a[3].y : (q == 2 ? a[3].z : a[3].w));
                        split_fast(e0, vh.x, vl.x); split_fast(e1, vh.y, vl.y); split_fast(e2, vh.z, vl.z); split_fast(e3, vh.w, vl.w);
                        dacc[q] = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, fmaf(e3, e3, dacc[q]))));
                        *reinterpret_cast<float4 *>(hi + qoff + q * 16) = vh;
                        *reinterpret_cast<float4 *>(lo + qoff + q * 16) = vl;
                    }
                }
            } else if (r < N) {
                const float *src = raw + (c * kGsChunkK + 4 * kh) * N + r;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int kk = 4 * kh + 8 * h;               // first k of this item inside the chunk
                    if (h < ksteps) {
                        float v0, v1, v2, v3;
                        if (kk + 3 < kleft) { v0 = src[0]; v1 = src[N]; v2 = src[2 * N]; v3 = src[3 * N]; }
                        else {
                            v0 = kk < kleft ? src[0] : 0.f; v1 = kk + 1 < kleft ? src[N] : 0.f;
                            v2 = kk + 2 < kleft ? src[2 * N] : 0.f; v3 = 0.f;
                        }
                        float4 vh, vl;
                        split_fast(v0, vh.x, vl.x); split_fast(v1, vh.y, vl.y); split_fast(v2, vh.z, vl.z); split_fast(v3, vh.w, vl.w);
                        dacc[0] = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, dacc[0]))));
                        *reinterpret_cast<float4 *>(hi + soff + h * 2 * kGsLbo) = vh;
                        *reinterpret_cast<float4 *>(lo + soff + h * 2 * kGsLbo) = vl;
                        src += 8 * N;
                    }
                }
            }
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                tc::fence_after_sync();
                const uint32_t a_hi = smem_u32(hi), a_lo = a_hi + p.tile_bytes;
                for (int kk = 0; kk < ksteps; ++kk) {
                    const uint32_t o = kk * 2 * kGsLbo;
                    const uint64_t dh = tc::smem_desc(a_hi + o, kGsLbo, kGsSbo), dl = tc::smem_desc(a_lo + o, kGsLbo, kGsSbo);
                    const uint32_t first = (c == 0 && kk == 0) ? 0u : 1u;
                    tc::mma_tf32(tmem + BN, dl, dh, idesc, first);        // cross terms: lo' hi + hi' lo
                    tc::mma_tf32(tmem + BN, dh, dl, idesc, 1u);
                    tc::mma_tf32(tmem, dh, dh, idesc, first);             // hi' hi
                }
                tc::mma_commit(&bar_done[s]);
                if (c == nchunks - 1) tc::mma_commit(bar_fin);
            }
        }
        // raw[rb] is free (every thread staged its last chunk before the barrier above): fetch the portfolio that uses it next
        if (use_tma && tid == 0) {
            const long long bn = b + (long long)p.nraw * b_step;
            if (bn < p.B) {
                fence_proxy_async();        // the in-place gamma scaling (generic proxy) is ordered before the bulk copy's writes
                mbar_expect_tx(&bar_raw[rb], nn_bytes);
                tma_bulk_g2s(raw, p.C + bn * (long long)N * N, nn_bytes, &bar_raw[rb]);
            }
        }
        // diag(Q) = column norms of A, combined over the k groups (the warm start of the solver reads this compact array)
        if (quads) {
            if (qk4 < 4) {
#pragma unroll
                for (int q = 0; q < 4; ++q) dpart[qk4 * 128 + 4 * rq + q] = dacc[q];
            }
        } else if (r < N) dpart[kh * 128 + r] = dacc[0];
        __syncthreads();
        if (tid < N) p.diag[b * N + tid] = quads ? (dpart[tid] + dpart[128 + tid]) + (dpart[256 + tid] + dpart[384 + tid]) : dpart[tid] + dpart[128 + tid];
        // epilogue: row m of Q = hi'hi + cross terms, 8 columns per step
        mbar_wait(bar_fin, (unsigned)(it & 1));
        tc::fence_after_sync();
        if (32 * (warp & 3) < N) {
            // stored transposed (Q is symmetric): a warp writes 32 consecutive floats of row n per instruction
            float *qcol = p.Q + (b * (long long)N + cbeg) * p.LDQ + m;
            const int ldq = p.LDQ;
#pragma unroll 1
            for (int c0 = cbeg; c0 < cend; c0 += 8, qcol += 8 * ldq) {
                float h8[8], x8[8];
                tmem_ld8(lane_addr + (uint32_t)c0, h8);
                tmem_ld8(lane_addr + (uint32_t)(BN + c0), x8);
                if (m < N) {
                    if (c0 + 8 <= N) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) qcol[j * ldq] = h8[j] + x8[j];
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (c0 + j < N) qcol[j * ldq] = h8[j] + x8[j];
                    }
                }
            }
        }
        tc::fence_before_sync();
        __syncthreads();          // accumulators read: the next portfolio's first MMA may overwrite them
        tc::fence_after_sync();
    }
    if (warp == 0) tc::tmem_dealloc(tmem, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------------
// (2) one warp per portfolio
// ------------------------------------------------------------------------------------------------------------------
constexpr int kQwFree = 32;            // largest reduced system (one row per lane)
constexpr int kQwLd = 33;              // row stride of the multiplier array (conflict-free rows and columns)
constexpr int kQwAssets = 4;           // assets per lane: n_assets <= 128

struct QwParams {
    const float *rets, *Q, *diag, *alpha;
    float u;
    long long B;
    int N, LDQ;
    float *w;
    int8_t *state;
    int32_t *status;
    int *fail_count, *fail_list, *work_count, *work_list, *bad_count;
};

// shared memory of one warp: mbarrier | multipliers L (32 x 33) | rets | reduced-row -> asset, -> slot | 32 row slots of Q
__host__ __device__ inline size_t qw_smem_bytes(int LDQ) { return 16 + (size_t)kQwFree * kQwLd * 4 + (size_t)LDQ * 4 + 256 + (size_t)kQwFree * LDQ * 4; }

// One elimination step of the reduced system, pivot row = lane K, with every register index a compile-time constant.
// The system sits at the END of the 32 x 32 frame (rows = lanes k0..31, columns = registers k0..31), so a system of any
// size runs the tail K = k0..31 of one unrolled sequence (entered through a switch) and executes exactly its own
// triangular share of shuffle + FMA pairs: no shifting, no per-group branches.
template <int K>
__device__ __forceinline__ void qw_step(float (&v)[kQwFree], float &y1, float &y2, float &dinv_me, int &lifted, float dmin,
                                        float *Lt, int lane) {
    float d = __shfl_sync(kFullMask, v[K], K);
    if (!(d > dmin)) { d = dmin; lifted = 1; }
    const float inv = recipFastT(d);
    const float l = lane > K ? v[K] * inv : 0.f;
    if (lane == K) dinv_me = inv;
    Lt[K * kQwLd + lane] = l;
#pragma unroll
    for (int j = K + 1; j < kQwFree; ++j) {
        const float rk = __shfl_sync(kFullMask, v[j], K);
        v[j] = fmaf(-l, rk, v[j]);
    }
    const float y1k = __shfl_sync(kFullMask, y1, K), y2k = __shfl_sync(kFullMask, y2, K);
    y1 = fmaf(-l, y1k, y1); y2 = fmaf(-l, y2k, y2);
}

__global__ void __launch_bounds__(32, 12) markowitz_qwarp_kernel(QwParams p) {
    extern __shared__ __align__(16) unsigned char qw_smem[];
    const long long b = blockIdx.x;
    const int lane = threadIdx.x, N = p.N, LDQ = p.LDQ;
    const float u = p.u;
    float *w_out = p.w + b * N;
    int8_t *st_out = p.state + b * N;
    // decided by the box alone (trivial_cases of markowitz_common.cuh)
    {
        const float cap = u * (float)N, tol = 8.f * Num<float>::eps() * (float)N;
        if (cap < 1.f - tol || !(u > 0.f)) {
            for (int i = lane; i < N; i += 32) { w_out[i] = Num<float>::inf() - Num<float>::inf(); st_out[i] = 0; }
            if (lane == 0) { p.status[b] = DDW_STATUS_INFEASIBLE; atomicAdd(p.bad_count, 1); }
            return;
        }
        if (cap <= 1.f + tol) {
            for (int i = lane; i < N; i += 32) { w_out[i] = u; st_out[i] = 1; }
            if (lane == 0) p.status[b] = DDW_STATUS_OK;
            return;
        }
    }
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(qw_smem);
    float *Lt = reinterpret_cast<float *>(qw_smem + 16);
    float *rs = Lt + kQwFree * kQwLd;
    int *idx = reinterpret_cast<int *>(rs + LDQ);
    int *sls = idx + kQwFree;
    float *rows = reinterpret_cast<float *>(sls + kQwFree);
    const unsigned row_bytes = (unsigned)LDQ * 4u;
    if (lane == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    const float *Qb = p.Q + b * (long long)N * LDQ;      // Q = A'A without the |alpha| I term, added where the diagonal is used
    const float aabs = fabsf(p.alpha[b]);
    const unsigned lt = (1u << lane) - 1u;
    float rr[kQwAssets], i2q[kQwAssets], wv[kQwAssets];
    int st[kQwAssets], slot[kQwAssets];
    bool in[kQwAssets];
    // ---- warm start: safeguarded Newton on the multiplier of the separable problem with diag(Q) (diag_init_warp)
    float dmax = 0.f;
    {
        float lo = Num<float>::inf(), hi = -Num<float>::inf(), sa = 0.f, sb = 0.f;
#pragma unroll
        for (int m = 0; m < kQwAssets; ++m) {
            const int i = lane + 32 * m;
            in[m] = i < N;
            rr[m] = 0.f; i2q[m] = 0.f; wv[m] = 0.f; st[m] = 2; slot[m] = -1;
            if (in[m]) {
                float q = p.diag[b * N + i] + aabs;
                rr[m] = p.rets[b * N + i];
                rs[i] = rr[m];
                dmax = fmaxf(dmax, q);
                q = q > Num<float>::tiny() ? q : Num<float>::tiny();
                i2q[m] = 0.5f / q;
                const float l = rr[m] - 2.f * q * u;
                lo = fminf(lo, l); hi = fmaxf(hi, rr[m]);
                sa = fmaf(rr[m], i2q[m], sa); sb += i2q[m];
            }
        }
        lo = warp_min(lo); hi = warp_max(hi); sa = warp_sum(sa); sb = warp_sum(sb); dmax = warp_max(dmax);
        float nu = clampT((sa - 1.f) / sb, lo, hi);
        float dxold = hi - lo, dx = dxold;
#pragma unroll 1
        for (int it = 0; it < 60; ++it) {
            float sum = 0.f, slope = 0.f;
#pragma unroll
            for (int m = 0; m < kQwAssets; ++m) {
                const float v = (rr[m] - nu) * i2q[m];
                sum += clampT(v, 0.f, u);
                slope += (v > 0.f && v < u) ? i2q[m] : 0.f;
            }
            sum = warp_sum(sum); slope = warp_sum(slope);
            if (absT(sum - 1.f) <= 128.f * Num<float>::eps()) break;
            if (sum > 1.f) lo = nu; else hi = nu;
            float nn = nu;
            bool ok = false;
            if (slope > 0.f) {
                nn = nu + (sum - 1.f) / slope;
                ok = nn > lo && nn < hi && absT(2.f * (nn - nu)) <= absT(dxold);
            }
            dxold = dx;
            if (ok) dx = nn - nu;
            else { dx = 0.5f * (hi - lo); nn = lo + dx; }
            if (nn == nu) break;
            nu = nn;
        }
#pragma unroll
        for (int m = 0; m < kQwAssets; ++m)
            if (in[m]) {
                const float v = (rr[m] - nu) * i2q[m];
                st[m] = v <= 0.f ? -1 : (v >= u ? 1 : 0);
            }
    }
    const float dmin = 4.f * Num<float>::eps() * dmax + Num<float>::tiny();
    const float tolw = 8.f * Num<float>::eps();
    int lifted = 0, converged = 0, bail = 0;
    unsigned phase = 0;
#pragma unroll 1
    for (int iter = 0; iter < kMaxPdas; ++iter) {
        // ---- (a) free / capped sets as ballots; position of every free asset in the reduced system
        unsigned mu[kQwAssets];
        int nF = 0, nU = 0, pos[kQwAssets];
#pragma unroll
        for (int m = 0; m < kQwAssets; ++m) {
            const unsigned mf = __ballot_sync(kFullMask, st[m] == 0);
            mu[m] = __ballot_sync(kFullMask, st[m] == 1);
            pos[m] = nF + __popc(mf & lt);
            nF += __popc(mf); nU += __popc(mu[m]);
        }
        if (nF + nU > kQwFree) { bail = 1; break; }
        const int k0 = kQwFree - nF;                       // the reduced system occupies lanes / registers k0 .. 31
        // ---- (b) rows of Q of the free and capped assets -> shared memory: one bulk copy per NEW asset (TMA), the slot of an
        // asset that left both sets is handed on.  The solver itself then never waits on global memory.
        {
            unsigned used = 0;
#pragma unroll
            for (int m = 0; m < kQwAssets; ++m) {
                const bool hold = st[m] == 0 || st[m] == 1;
                used |= __reduce_or_sync(kFullMask, (hold && slot[m] >= 0) ? 1u << slot[m] : 0u);
                if (!hold) slot[m] = -1;
            }
            const unsigned freem = ~used;
            int nnew = 0;
            unsigned needm[kQwAssets];
#pragma unroll
            for (int m = 0; m < kQwAssets; ++m) {
                const bool need = (st[m] == 0 || st[m] == 1) && slot[m] < 0;
                needm[m] = __ballot_sync(kFullMask, need);
                if (need) slot[m] = (int)__fns(freem, 0, nnew + __popc(needm[m] & lt) + 1);
                nnew += __popc(needm[m]);
            }
            __syncwarp();                                  // every read of the slots that are handed on has been issued
            if (nnew > 0) {
                if (lane == 0) mbar_expect_tx(mbar, (unsigned)nnew * row_bytes);
                __syncwarp();
#pragma unroll
                for (int m = 0; m < kQwAssets; ++m)
                    if ((needm[m] >> lane) & 1u) tma_bulk_g2s(rows + slot[m] * LDQ, Qb + (lane + 32 * m) * LDQ, row_bytes, mbar);
            }
#pragma unroll
            for (int m = 0; m < kQwAssets; ++m)
                if (st[m] == 0) { idx[pos[m]] = lane + 32 * m; sls[pos[m]] = slot[m]; }
            __syncwarp();
            if (nnew > 0) { mbar_wait(mbar, phase); phase ^= 1u; }
        }
        const bool live = lane >= k0;
        const int a = live ? idx[lane - k0] : 0;           // asset and row slot of reduced row `lane`
        const int sa = live ? sls[lane - k0] : 0;
        const float c = 1.f - u * (float)nU;
        float nu_mult = 0.f, x = 0.f;
        if (nF > 0) {
            // ---- (c) row `lane` of Q_FF (+ |alpha| on the diagonal) by symmetry from row slot j, column a: one row per
            // instruction, no bank conflicts; right-hand sides r_F - 2 u Q_FU 1 and 1
            float v[kQwFree];
#pragma unroll
            for (int j0 = 0; j0 < kQwFree; j0 += 4) {
                if (j0 + 3 >= k0) {
#pragma unroll
                    for (int j = j0; j < j0 + 4; ++j) {
                        const int sj = __shfl_sync(kFullMask, sa, j);
                        v[j] = (live && j >= k0) ? rows[sj * LDQ + a] : 0.f;
                        if (j == lane) v[j] += aabs;
                    }
                } else {
#pragma unroll
                    for (int j = j0; j < j0 + 4; ++j) v[j] = 0.f;
                }
            }
            float y1 = live ? rs[a] : 0.f, y2 = live ? 1.f : 0.f;
            if (nU > 0) {
                float acc = 0.f;
#pragma unroll
                for (int m = 0; m < kQwAssets; ++m) {
                    unsigned mk = mu[m];
                    while (mk) {
                        const int bit = __ffs(mk) - 1;
                        mk &= mk - 1;
                        if (live) acc += rows[sa * LDQ + 32 * m + bit];
                    }
                }
                y1 -= 2.f * u * acc;
            }
            // ---- (d) elimination without pivoting (LDL'), steps k0 .. 31 of the unrolled sequence
            float dinv_me = 0.f;
#define QW_CASE(K) case K: qw_step<K>(v, y1, y2, dinv_me, lifted, dmin, Lt, lane);
            switch (k0) {
                QW_CASE(0) QW_CASE(1) QW_CASE(2) QW_CASE(3) QW_CASE(4) QW_CASE(5) QW_CASE(6) QW_CASE(7)
                QW_CASE(8) QW_CASE(9) QW_CASE(10) QW_CASE(11) QW_CASE(12) QW_CASE(13) QW_CASE(14) QW_CASE(15)
                QW_CASE(16) QW_CASE(17) QW_CASE(18) QW_CASE(19) QW_CASE(20) QW_CASE(21) QW_CASE(22) QW_CASE(23)
                QW_CASE(24) QW_CASE(25) QW_CASE(26) QW_CASE(27) QW_CASE(28) QW_CASE(29) QW_CASE(30) QW_CASE(31)
                default: break;
            }
#undef QW_CASE
            // ---- (e) multiplier of the budget constraint, then L' x = D^-1 z column by column
            const float s12 = warp_sum(y1 * y2 * dinv_me), s22 = warp_sum(y2 * y2 * dinv_me);
            const float nu = (s12 - 2.f * c) / s22;
            x = 0.5f * (y1 - nu * y2) * dinv_me;
            __syncwarp();
#pragma unroll 1
            for (int i = kQwFree - 1; i > k0; --i) {
                const float xi = __shfl_sync(kFullMask, x, i);
                if (live && lane < i) x = fmaf(-Lt[lane * kQwLd + i], xi, x);
            }
            nu_mult = nu;
        }
        // ---- (f) weights in asset order, gradient 2 (Q + |alpha| I) w - r + nu from the row slots
        float g[kQwAssets];
#pragma unroll
        for (int m = 0; m < kQwAssets; ++m) {
            const float xf = __shfl_sync(kFullMask, x, (k0 + pos[m]) & 31);
            wv[m] = st[m] == 0 ? xf : (st[m] == 1 ? u : 0.f);
            g[m] = aabs * wv[m];
        }
#pragma unroll 4
        for (int j = k0; j < kQwFree; ++j) {
            const int sj = __shfl_sync(kFullMask, sa, j);
            const float xj = __shfl_sync(kFullMask, x, j);
            const float *row = rows + sj * LDQ + lane;
#pragma unroll
            for (int m = 0; m < kQwAssets; ++m)
                if (in[m]) g[m] = fmaf(row[32 * m], xj, g[m]);
        }
        if (nU > 0) {
#pragma unroll
            for (int m2 = 0; m2 < kQwAssets; ++m2) {
                unsigned mk = mu[m2];
                while (mk) {
                    const int bit = __ffs(mk) - 1;
                    mk &= mk - 1;
                    const int sc = __shfl_sync(kFullMask, slot[m2], bit);
                    const float *row = rows + sc * LDQ + lane;
#pragma unroll
                    for (int m = 0; m < kQwAssets; ++m)
                        if (in[m]) g[m] = fmaf(row[32 * m], u, g[m]);
                }
            }
        }
        int changed = 0;
        if (nF > 0) {
#pragma unroll
            for (int m = 0; m < kQwAssets; ++m) {
                if (!in[m]) continue;
                const float gi = 2.f * g[m] - rr[m] + nu_mult;
                int sn = st[m];
                if (st[m] == 0) {
                    if (wv[m] < -tolw) sn = -1;
                    else if (wv[m] > u + tolw) sn = 1;
                } else if (st[m] < 0) { if (gi < 0.f) sn = 0; }
                else { if (gi > 0.f) sn = 0; }
                if (sn != st[m]) { st[m] = sn; changed = 1; }
            }
        } else {
            // no free asset: optimal iff max_L(-h) <= min_U(-h) and u |U| = 1 (same rule as the other kernels)
            float hl = Num<float>::inf(), hu = Num<float>::inf();
            int il = 0x7fffffff, iu = 0x7fffffff;
#pragma unroll
            for (int m = 0; m < kQwAssets; ++m) {
                if (!in[m]) continue;
                const float h = 2.f * g[m] - rr[m];
                const int i = lane + 32 * m;
                if (st[m] < 0) { if (h < hl) { hl = h; il = i; } }
                else if (st[m] > 0) { if (-h < hu) { hu = -h; iu = i; } }
            }
            warp_argmin(hl, il); warp_argmin(hu, iu);
            const float ctol = 4.f * Num<float>::eps() * (float)N;
            int fl = -1, fu = -1;
            if (c > ctol && il < N) fl = il;
            else if (c < -ctol && iu < N) fu = iu;
            else if (il < N && iu < N && -hl > hu) { fl = il; fu = iu; }
#pragma unroll
            for (int m = 0; m < kQwAssets; ++m) {
                const int i = lane + 32 * m;
                if (i == fl || i == fu) { st[m] = 0; changed = 1; }
            }
        }
        if (!__any_sync(kFullMask, changed)) { converged = 1; break; }
    }
    if (bail) {
        if (lane == 0) { const int sl = atomicAdd(p.work_count, 1); p.work_list[sl] = (int)b; }
        return;
    }
    int bad = 0;
#pragma unroll
    for (int m = 0; m < kQwAssets; ++m) {
        if (!in[m]) continue;
        const int i = lane + 32 * m;
        int s8 = st[m];
        const float wi = s8 == 0 ? clampT(wv[m], 0.f, u) : (s8 > 0 ? u : 0.f);
        if (s8 == 0) { bad |= !(absT(wv[m]) < Num<float>::inf()); s8 = wi <= 0.f ? -1 : (wi >= u ? 1 : 0); }   // tie rule, see write_result
        bad |= !(absT(rr[m]) < Num<float>::inf());
        w_out[i] = wi;
        st_out[i] = (int8_t)s8;
    }
    const int nonfinite = __any_sync(kFullMask, bad);
    if (lane == 0) {
        int s = lifted ? DDW_STATUS_REGULARIZED : 0;
        if (nonfinite) { s |= DDW_STATUS_INEXACT; atomicAdd(p.bad_count, 1); }
        else if (!converged) {
            s |= DDW_STATUS_FALLBACK | DDW_STATUS_INEXACT;
            const int sl = atomicAdd(p.fail_count, 1);
            p.fail_list[sl] = (int)b;
        }
        p.status[b] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
bool fwd_qwarp_eligible(int N, int tsize) {
    static const bool on = getenv("DDW_QWARP") != nullptr;      // work in progress: opt-in until it beats the one-CTA-per-portfolio kernel
    return on && tsize == 4 && N >= 8 && N <= 128;
}
int fwd_qwarp_ldq(int N) { return (N + 7) / 8 * 8; }     // rows of the scratch start on 32-byte sectors
// Q (B x N x LDQ) followed by diag(Q) (B x N)
size_t fwd_qwarp_scratch_bytes(long long B, int N) { return (size_t)B * N * (fwd_qwarp_ldq(N) + 1) * sizeof(float); }

int launch_fwd_qwarp(const MkFwdParams<float> &p, float *Q, cudaStream_t st) {
    static DeviceOnce configured;
    const int dev = current_device();
    float *diag = Q + (size_t)p.B * p.N * fwd_qwarp_ldq(p.N);
    int smem_sm = 0, nsm = 0;
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const GsLayout l = gs_layout(p.N, smem_sm);
    if (configured.needs(dev)) {
        if (cudaFuncSetAttribute(markowitz_gram_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, device_smem_limit()) != cudaSuccess ||
            cudaFuncSetAttribute(markowitz_qwarp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qw_smem_bytes(128)) != cudaSuccess)
            return check_launch("cudaFuncSetAttribute(markowitz_gram_small_kernel / markowitz_qwarp_kernel)");
        configured.mark(dev);
    }
    GsParams g;
    g.C = p.C; g.gamma = p.gamma; g.alpha = p.alpha; g.Q = Q; g.diag = diag; g.B = p.B; g.Bmod = p.Bmod; g.b0 = p.b0;
    g.N = p.N; g.LDQ = fwd_qwarp_ldq(p.N); g.BN = l.BN; g.nraw = l.nraw;
    g.raw_bytes = l.raw_bytes; g.tile_bytes = l.tile_bytes; g.tmem_cols = l.tmem_cols;
    const int per_sm = 2 * ((size_t)l.total + 1024) <= (size_t)smem_sm ? 2 : 1;
    const long long grid = (long long)per_sm * (nsm > 0 ? nsm : 148);
    markowitz_gram_small_kernel<<<(unsigned)(p.B < grid ? p.B : grid), kGsThreads, l.total, st>>>(g);
    int rc = check_launch("markowitz_gram_small_kernel");
    if (rc) return rc;
    QwParams q;
    q.rets = p.rets; q.Q = Q; q.diag = diag; q.alpha = p.alpha; q.u = p.u; q.B = p.B; q.N = p.N; q.LDQ = g.LDQ; q.w = p.w; q.state = p.state; q.status = p.status;
    q.fail_count = p.fail_count; q.fail_list = p.fail_list; q.work_count = p.work_count; q.work_list = p.work_list;
    q.bad_count = p.bad_count;
    markowitz_qwarp_kernel<<<(unsigned)p.B, 32, qw_smem_bytes(g.LDQ), st>>>(q);
    return check_launch("markowitz_qwarp_kernel");
}

}  // namespace ddw
