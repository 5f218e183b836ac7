// markowitz_qwarp.cu -- float32 NumericalMarkowitz forward for n_assets <= 128 as two kernels:
//
//   (1) markowitz_gram_small_kernel   Q = A'A of every portfolio on the 5th-generation tensor cores
//       (tcgen05.mma.kind::tf32 in the 3xTF32 split, accumulators in TMEM).  A = gamma-tiled covmat_sqrt
//       (reference deepdow/layers/allocate.py:268-273) arrives by ONE bulk TMA copy per portfolio (4 N^2 bytes,
//       2-3 portfolios ahead), is split into TF32 hi/lo parts chunk by chunk into the tensor core's K-major SWIZZLE_128B
//       layout, and Q leaves through tcgen05.ld as coalesced stores into a scratch in the workspace (B x N x LDQ floats).
//   (2) markowitz_qwarp_kernel        the primal-dual active-set iteration of markowitz_sparse.cuh with ONE WARP per
//       portfolio and no shared copy of A or Q: the reduced KKT system of the free assets (<= 32) lives in registers,
//       one row per lane, and is eliminated with warp shuffles; the few rows of Q it needs arrive by one bulk TMA copy per
//       asset into 32 row slots of the warp's shared memory.
//
// Why: the one-CTA-per-portfolio kernel keeps 40 KB of A resident per portfolio (4 portfolios per SM) while ~60% of its
// time is spent in phases that never touch A, and spends ~37 k warp instructions per portfolio, most of them index
// arithmetic and barriers around ~2.5 k warp instructions of useful FMAs.  With Q formed by the tensor core the solver
// needs only |F| rows of Q per iteration; one warp per portfolio removes every block barrier and lets 14 portfolios
// share an SM.  A probe kernel (0) first decides whether the batch is in the diversified regime (then everything goes to the dense
// kernel at once).  Free sets that outgrow the 32-row frame are trimmed; what still does not fit goes to the one-CTA-per-portfolio
// kernel in list mode, beyond 256 of those to the dense kernel (work_list); portfolios that do not converge in kMaxPdas iterations
// go to the interior-point kernel (fail_list) -- the protocol of the one-CTA-per-portfolio kernel, which stays the float64 path and
// the float32 path for batches below 1024 portfolios.
#include <stdlib.h>
#include "markowitz_common.cuh"
#include "support_kernels.cuh"
#include "tc_gemm.cuh"

namespace ddw {

// probe word (section 0 below): bits 0-15 portfolios that overflow the frame, bits 16-30 portfolios probed, bit 31 "the batch is
// the dense kernel's"
__device__ __forceinline__ bool probe_says_dense(const int *word) { return *reinterpret_cast<const volatile int *>(word) < 0; }

// ------------------------------------------------------------------------------------------------------------------
// (1) Gram kernel
// ------------------------------------------------------------------------------------------------------------------
constexpr int kGsEpiWarps = 8;                      // epilogue: TMEM lane quadrant = warp % 4, column half = warp / 4
constexpr int kGsStageWarps = 4;                    // staging: split A into TF32 hi / lo tiles; warps per group
constexpr int kGsStageGroups = 2;                   // group g stages the chunks c % 2 == g of every portfolio
constexpr int kGsStageThreads = kGsStageWarps * 32;
constexpr int kGsStageAll = kGsStageGroups * kGsStageThreads;
constexpr int kGsMmaWarp = kGsEpiWarps + kGsStageGroups * kGsStageWarps, kGsTmaWarp = kGsMmaWarp + 1;
constexpr int kGsPollWarp = kGsTmaWarp + 1;         // watches the mbarriers for the MMA thread (see there)
constexpr int kGsThreads = (kGsPollWarp + 1) * 32;  // 608
constexpr int kGsRing = 4;                          // staged chunks in flight
constexpr int kGsMaxRaw = 3;                        // raw A buffers (TMA prefetch depth)
constexpr int kGsChunkK = 32;                       // k values per staged chunk = one 128-byte swizzle row = 4 MMA k-steps
constexpr int kGsRowBytes = 128, kGsSbo = 8 * kGsRowBytes;   // K-major SWIZZLE_128B: 8-row atoms of 1024 B
constexpr int kGsHead = 1024;                       // barriers and flags (keeps the tiles behind it 1024-byte aligned)

struct GsLayout {
    int BN;                 // MMA N extent = n_assets rounded up to 16
    int nraw;               // raw A buffers
    unsigned raw_bytes;     // one raw buffer, rounded to 128 B
    unsigned tile_bytes;    // one hi or lo tile: BN rows x 32 k
    unsigned ring_bytes;    // kGsRing stages of {hi, lo} + the over-read pad
    unsigned total;         // dynamic shared memory of the kernel
    unsigned tmem_cols;
};
static inline GsLayout gs_layout(int N, int smem_limit) {
    GsLayout l;
    l.BN = (N + 15) / 16 * 16;
    l.raw_bytes = (unsigned)(((size_t)N * N * 4 + 127) / 128 * 128);
    l.tile_bytes = (unsigned)(l.BN / 8) * kGsSbo;
    // the M = 128 operand reads 16 row groups of every tile: (16 - BN/8) groups past its end.  They land in the next tile of
    // the ring (finite or not, they only reach output rows >= BN, which nobody reads); the pad keeps the last tile's
    // over-read inside the allocation
    l.ring_bytes = 2 * kGsRing * l.tile_bytes + (unsigned)(16 - l.BN / 8) * kGsSbo;
    const unsigned fixed = 1024 /* alignment slack */ + kGsHead + l.ring_bytes;
    int nraw = ((unsigned)smem_limit - fixed) / l.raw_bytes;
    l.nraw = nraw > kGsMaxRaw ? kGsMaxRaw : (nraw < 1 ? 1 : nraw);
    l.total = fixed + l.nraw * l.raw_bytes;
    const unsigned need = 4u * (unsigned)l.BN;          // two accumulator buffers x {hi'hi, cross terms}
    l.tmem_cols = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512)));
    return l;
}

struct GsParams {
    const float *C, *gamma, *alpha;
    float *Q;
    long long B, Bmod, b0;
    int N, LDQ, BN, nraw;
    unsigned raw_bytes, tile_bytes, ring_bytes, tmem_cols;
    const int *skip;        // probe flag: the batch is the dense kernel's
};

// Polling wait (mbarrier.test_wait): the role warps hand work to each other every few hundred cycles, and the suspending
// try_wait of common.cuh was measured to add about a microsecond per hand-over here (an empty pipeline skeleton took
// 6.3 k cycles per portfolio).  Bounded like mbar_wait: a barrier that never completes traps instead of hanging the device.
__device__ __forceinline__ void mbar_poll(unsigned long long *bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    for (int spin = 0; spin < (1 << 28); ++spin) {
        unsigned ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void stage_group_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kGsStageAll) : "memory"); }
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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
__device__ __forceinline__ void gamma_tiling_batched(float *As, int N, const float *__restrict__ gamma, long long gflat0, long long Bmod, int tid) {
    const unsigned uB = (unsigned)Bmod;
    if ((N & 3) == 0 && (uB & 3u) == 0u) {
        unsigned gi = (unsigned)((gflat0 + 4 * tid) % Bmod);
        const unsigned gstep = (unsigned)((4 * kGsStageAll) % Bmod);
        const int NN = N * N;
        for (int e0 = 4 * tid; e0 < NN; e0 += 16 * kGsStageAll) {
            float4 g[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (e0 + q * 4 * kGsStageAll < NN) g[q] = __ldg(reinterpret_cast<const float4 *>(gamma + gi));
                gi += gstep; if (gi >= uB) gi -= uB;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int e = e0 + q * 4 * kGsStageAll;
                if (e < NN) {
                    float4 a = *reinterpret_cast<float4 *>(As + e);
                    a.x *= g[q].x; a.y *= g[q].y; a.z *= g[q].z; a.w *= g[q].w;
                    *reinterpret_cast<float4 *>(As + e) = a;
                }
            }
        }
    } else {
        unsigned gi = (unsigned)((gflat0 + tid) % Bmod);
        const unsigned gstep = (unsigned)(kGsStageAll % Bmod);
        for (int e = tid; e < N * N; e += kGsStageAll) {
            As[e] *= __ldg(gamma + gi);
            gi += gstep; if (gi >= uB) gi -= uB;
        }
    }
}

// Optional role timing (build with -DDDW_GS_TIMING, tools/gs_timing.py): one thread per role adds the cycles it spent in
// each of its waits / work sections to a global table.  Compiled out of the product build.
#ifdef DDW_GS_TIMING
__device__ unsigned long long ddw_gs_cycles[16];
__device__ long long ddw_gs_trace[4096];      // [role 0..3][portfolio 0..3][chunk 0..7][begin, end] of CTA 0
#define GS_TR(role, it, c, e) do { if (blockIdx.x == 0 && (it) < 4) ddw_gs_trace[(((role) * 4 + (int)(it)) * 8 + (c)) * 2 + (e)] = clock64(); } while (0)
// per-instruction timestamps of the MMA thread: portfolio 2 of CTA 0, [chunk][14] behind the role table
#define GS_TI(it, c, i) do { if (blockIdx.x == 0 && (it) == 2) ddw_gs_trace[256 + (c) * 16 + (i)] = clock64(); } while (0)
#define GS_T0 long long gs_t0 = clock64();
#define GS_T(id, cond) do { if (cond) { const long long gs_t1 = clock64(); atomicAdd(&ddw_gs_cycles[id], (unsigned long long)(gs_t1 - gs_t0)); gs_t0 = gs_t1; } } while (0)
#else
#define GS_T0
#define GS_T(id, cond) do { } while (0)
#define GS_TR(role, it, c, e) do { } while (0)
#define GS_TI(it, c, i) do { } while (0)
#endif

// Warp-specialised, one persistent CTA per SM:
//   TMA warp       bulk copies of A (one per portfolio) into a ring of raw buffers, kGsMaxRaw portfolios ahead
//   staging warps  gamma tiling, TF32 hi / lo split into a ring of kGsRing operand stages (K-major, 128-byte swizzle)
//   MMA warp       one thread issues the 3xTF32 tcgen05.mma of every staged chunk into one of two TMEM accumulator buffers
//   epilogue warps tcgen05.ld of the finished buffer -> coalesced stores of Q
// The roles meet only through mbarriers (full / empty per ring slot), so the epilogue of portfolio p, the MMAs of p + 1 and
// the staging of p + 2 overlap.
__global__ void __launch_bounds__(kGsThreads, 1) markowitz_gram_small_kernel(GsParams p) {
    extern __shared__ unsigned char gs_smem[];
    if (probe_says_dense(p.skip)) return;     // diversified batch (probe kernel): nothing to do here
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, BN = p.BN;
    // 1 KB alignment by an offset from the shared array itself: the pointers keep their address space (LDS / STS, not generic)
    unsigned char *base = gs_smem + ((1024u - (smem_u32(gs_smem) & 1023u)) & 1023u);
    unsigned long long *raw_full = reinterpret_cast<unsigned long long *>(base);         // [kGsMaxRaw]
    unsigned long long *raw_empty = raw_full + kGsMaxRaw;                                // [kGsMaxRaw]
    unsigned long long *ring_full = raw_empty + kGsMaxRaw;                               // [kGsRing]
    unsigned long long *ring_empty = ring_full + kGsRing;                                // [kGsRing / 2]: one per pair of stages
    unsigned long long *acc_full = ring_empty + kGsRing;                                 // [2]
    unsigned long long *acc_empty = acc_full + 2;                                        // [2]
    uint32_t &tmem_slot = *reinterpret_cast<uint32_t *>(base + 192);
    unsigned *ready = reinterpret_cast<unsigned *>(base + 256);                          // chunks the MMA thread may issue (poller -> MMA)
    unsigned char *ring = base + kGsHead;
    const unsigned stage_bytes = 2 * p.tile_bytes;
    unsigned char *raw0 = ring + p.ring_bytes;
    const unsigned nn_bytes = (unsigned)(N * N * 4);
    const bool use_tma = (nn_bytes & 15u) == 0u && (reinterpret_cast<unsigned long long>(p.C) & 15ull) == 0ull;

    if (tid == 0) {
        for (int i = 0; i < kGsMaxRaw; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], kGsStageGroups * kGsStageWarps); }
        for (int i = 0; i < kGsRing; ++i) { mbar_init(&ring_full[i], kGsStageWarps); mbar_init(&ring_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kGsEpiWarps); }
        *ready = 0u;
        fence_mbar_init();
    }
    // rows >= n_assets of the tiles are never written again: zero the whole ring once
    for (unsigned e = tid * 16u; e < p.ring_bytes; e += kGsThreads * 16u)
        *reinterpret_cast<float4 *>(ring + e) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == kGsMmaWarp) tc::tmem_alloc(&tmem_slot, p.tmem_cols);
    // gamma_sqrt the same for the whole batch (the usual case: a learned scalar, repeated): Q = gamma^2 A'A, scaled in the
    // epilogue, and the in-place tiling pass over A is skipped
    int gsame = p.Bmod <= (1 << 16);
    if (gsame) {
        const float g0 = __ldg(p.gamma);
        for (long long i = tid; i < p.Bmod; i += kGsThreads) gsame &= __ldg(p.gamma + i) == g0;
    }
    tc::fence_before_sync();
    const bool gunif = __syncthreads_and(gsame) != 0;
    tc::fence_after_sync();
    const float gsq = gunif ? __ldg(p.gamma) * __ldg(p.gamma) : 1.f;
    const uint32_t tmem = tmem_slot;
    const long long b_first = blockIdx.x, b_step = gridDim.x;
    const int nchunks = (N + kGsChunkK - 1) / kGsChunkK;

    if (warp == kGsTmaWarp) {
        // ---------------- TMA producer
        if (use_tma) {
            long long it = 0;
            for (long long b = b_first; b < p.B; b += b_step, ++it) {
                if (lane == 0) {
                    GS_T0
                    const int buf = (int)(it % p.nraw);
                    const long long use = it / p.nraw;
                    if (use > 0) mbar_wait(&raw_empty[buf], (unsigned)((use - 1) & 1));
                    GS_T(0, true);
                    mbar_expect_tx(&raw_full[buf], nn_bytes);
                    tma_bulk_g2s(raw0 + (size_t)buf * p.raw_bytes, p.C + b * (long long)N * N, nn_bytes, &raw_full[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp == kGsMmaWarp) {
        // ---------------- MMA issuer
        const uint32_t idesc = tc::idesc_tf32(128, BN);
        // K-major SWIZZLE_128B descriptor: start address >> 4 | LBO field 1 | SBO = 1024 B | version 1 | layout type 2.
        // Descriptors of the stages, tiles and k-steps differ only in the 14-bit address field: one add per operand and MMA.
        const uint64_t desc0 = (uint64_t)((smem_u32(ring) >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(kGsSbo >> 4) << 32) |
                               (1ull << 46) | (2ull << 61);
        const uint64_t dstage = stage_bytes >> 4, dtile = p.tile_bytes >> 4;
        // A tcgen05.commit is handed to the tensor pipe, and every later mbarrier operation of the SAME thread was measured to
        // wait until that commit has fired, i.e. until the MMAs before it are done (~600 idle cycles after every chunk: the
        // pipe drained while this thread sat in a poll that was already satisfied).  So this thread never touches an mbarrier:
        // the poller warp waits for "accumulator free" and "stage full" and publishes the number of chunks that may be issued.
        unsigned g = 0, have = 0;
        long long it = 0;
        for (long long b = b_first; b < p.B; b += b_step, ++it) {
            if (lane == 0) {
                const unsigned tb = (unsigned)(it & 1);
                GS_T0
                const uint32_t acc = tmem + tb * 2u * (uint32_t)BN;
                for (int c = 0; c < nchunks; ++c, ++g) {
                    const unsigned s = (unsigned)c;                  // every portfolio starts at stage 0 (at most kGsRing chunks)
                    // `have` was refreshed BEFORE the previous chunk's commit (below): in steady state the stagers are ahead and
                    // this thread goes from a commit straight to the next MMAs without touching shared memory in between
                    for (int spin = 0; have <= g; ++spin) {
                        asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(have) : "r"(smem_u32(ready)) : "memory");
                        if (spin > (1 << 28)) __trap();
                    }
                    GS_TI(it, c, 14);
                    if (c == 0) tc::fence_after_sync();        // the epilogue's tcgen05.ld of this buffer are ordered before the MMAs
                    GS_TI(it, c, 15);
                    GS_T(2, true);
                    GS_TR(2, it, c, 0);
                    uint64_t dh = desc0 + s * dstage, dl = dh + dtile;
                    const int ksteps = min(4, (N - c * kGsChunkK + 7) >> 3);
                    for (int kk = 0; kk < ksteps; ++kk, dh += 2, dl += 2) {      // 8 k values = 32 B along the swizzle row
                        const uint32_t first = (c == 0 && kk == 0) ? 0u : 1u;
                        tc::mma_tf32(acc + BN, dl, dh, idesc, first);        // cross terms: lo' hi + hi' lo
                        GS_TI(it, c, 3 * kk);
                        tc::mma_tf32(acc + BN, dh, dl, idesc, 1u);
                        GS_TI(it, c, 3 * kk + 1);
                        tc::mma_tf32(acc, dh, dh, idesc, first);             // hi' hi
                        GS_TI(it, c, 3 * kk + 2);
                    }
                    // one commit per PAIR of stages: a tcgen05.commit holds this thread until the MMAs before it are done
                    // (~1.2 k cycles each, measured: five commits per portfolio were the period of the whole kernel)
                    // any shared-memory access of this thread queues behind a pending tcgen05.commit (measured: ~300 idle cycles
                    // after every commit, 1.1 k after the two at a portfolio boundary): read the flag first
                    asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(have) : "r"(smem_u32(ready)) : "memory");
                    if ((c & 1) || c == nchunks - 1) tc::mma_commit(&ring_empty[c >> 1]);
                    GS_TI(it, c, 12);
                    if (c == nchunks - 1) tc::mma_commit(&acc_full[tb]);
                    GS_TI(it, c, 13);
                    GS_T(3, true);
                    GS_TR(2, it, c, 1);
                }
            }
            __syncwarp();
        }
    } else if (warp == kGsPollWarp) {
        // ---------------- poller: mbarrier waits on behalf of the MMA thread
        unsigned g = 0;
        long long it = 0;
        for (long long b = b_first; b < p.B; b += b_step, ++it) {
            if (lane == 0) {
                const unsigned tb = (unsigned)(it & 1);
                const long long ut = it >> 1;
                if (ut > 0) mbar_poll(&acc_empty[tb], (unsigned)((ut - 1) & 1));      // the epilogue has read this buffer
                for (int c = 0; c < nchunks; ++c, ++g) {
                    mbar_poll(&ring_full[c], (unsigned)(it & 1));                     // (the stagers' fence.proxy.async made their stores visible to the MMA)
                    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(ready)), "r"(g + 1u) : "memory");
                }
            }
            __syncwarp();
        }
    } else if (warp >= kGsEpiWarps) {
        // ---------------- staging
        const int sta = tid - kGsEpiWarps * 32;                  // index among all staging threads (gamma pass, release)
        const int grp = sta / kGsStageThreads, st = sta % kGsStageThreads;
        // Element (row r, k) of a tile lives at (r / 8) * 1024 + (r % 8) * 128 + (((k / 4) ^ (r % 8)) * 16) + (k % 4) * 4.
        // Fast path (n_assets % 4 == 0): an item = (quad of operand rows 4 rq .. 4 rq + 3, k group k4 of the chunk): four 16-byte
        // loads of A (one per k), split, four 16-byte stores per tile (one per row); the 4 x 4 transpose is register naming.
        // Eight consecutive items = {k4 & 3} x {rq & 1}, which spreads a quarter-warp's stores over all eight 16-byte columns.
        // Generic path: thread = operand row, scalar loads.
        const bool quads = (N & 3) == 0;
        const int N4 = N >> 2, nitems = 16 * ((N4 + 1) >> 1);
        unsigned g = 0;
        long long it = 0;
        for (long long b = b_first; b < p.B; b += b_step, ++it) {
            const int buf = (int)(it % p.nraw);
            const long long use = it / p.nraw;
            float *raw = reinterpret_cast<float *>(raw0 + (size_t)buf * p.raw_bytes);
            GS_T0
            if (use_tma) mbar_poll(&raw_full[buf], (unsigned)(use & 1));
            else {
                const float *Cb = p.C + b * (long long)N * N;
                stage_group_sync();                // every chunk of the previous portfolio has been read
                for (int e = sta; e < N * N; e += kGsStageAll) raw[e] = Cb[e];
                stage_group_sync();
            }
            GS_T(5, sta == 0);
            if (!gunif) {
                gamma_tiling_batched(raw, N, p.gamma, (p.b0 + b) * (long long)N * N, p.Bmod, sta);
                stage_group_sync();
            }
            GS_T(6, sta == 0);
            for (int c = 0; c < nchunks; ++c, ++g) {
                if ((c & 1) != grp) continue;                        // the other group's chunk
                const unsigned s = (unsigned)c;
                unsigned char *hi = ring + s * stage_bytes, *lo = hi + p.tile_bytes;
                const int kleft = N - c * kGsChunkK;                 // k values of this chunk that exist
                const int kgroups = 2 * min(4, (kleft + 7) >> 3);    // 4-wide k groups the MMAs of this chunk read
                if (it > 0) mbar_poll(&ring_empty[c >> 1], (unsigned)((it - 1) & 1));   // the previous portfolio's MMAs on this pair of stages have drained
                GS_T(7, sta == 0);
                if (st == 0) GS_TR(grp, it, c, 0);
                if (quads) {
                    for (int e = st; e < nitems; e += kGsStageThreads) {
                        const int w = e & 15, k4 = (w & 3) + 4 * (w >> 3), rq = 2 * (e >> 4) + ((w >> 2) & 1);
                        if (rq >= N4 || k4 >= kgroups) continue;
                        float4 a[4];
                        if (4 * k4 < kleft) {
                            const float *src = raw + (c * kGsChunkK + 4 * k4) * N + 4 * rq;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) a[kk] = *reinterpret_cast<const float4 *>(src + kk * N);
                        } else {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) a[kk] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        const unsigned rbase = (unsigned)((rq >> 1) * kGsSbo + (rq & 1) * 4 * kGsRowBytes);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float4 vh, vl;
                            const float e0 = q == 0 ? a[0].x : (q == 1 ? a[0].y : (q == 2 ? a[0].z : a[0].w));
                            const float e1 = q == 0 ? a[1].x : (q == 1 ? a[1].y : (q == 2 ? a[1].z : a[1].w));
                            const float e2 = q == 0 ? a[2].x : (q == 1 ? a[2].y : (q == 2 ? a[2].z : a[2].w));
                            const float e3 = q == 0 ? a[3].x : (q == 1 ? a[3].y : (q == 2 ? a[3].z : a[3].w));
                            split_fast(e0, vh.x, vl.x); split_fast(e1, vh.y, vl.y); split_fast(e2, vh.z, vl.z); split_fast(e3, vh.w, vl.w);
                            const unsigned r7 = (unsigned)(4 * (rq & 1) + q);
                            const unsigned off = rbase + q * kGsRowBytes + ((unsigned)k4 ^ r7) * 16u;
                            *reinterpret_cast<float4 *>(hi + off) = vh;
                            *reinterpret_cast<float4 *>(lo + off) = vl;
                        }
                    }
                } else if (st < N) {
                    const float *src = raw + c * kGsChunkK * N + st;
                    const unsigned rbase = (unsigned)((st >> 3) * kGsSbo + (st & 7) * kGsRowBytes);
                    for (int k4 = 0; k4 < kgroups; ++k4, src += 4 * N) {
                        const int kk = 4 * k4;                       // first k of this item inside the chunk
                        const float v0 = kk < kleft ? src[0] : 0.f, v1 = kk + 1 < kleft ? src[N] : 0.f;
                        const float v2 = kk + 2 < kleft ? src[2 * N] : 0.f, v3 = kk + 3 < kleft ? src[3 * N] : 0.f;
                        float4 vh, vl;
                        split_fast(v0, vh.x, vl.x); split_fast(v1, vh.y, vl.y); split_fast(v2, vh.z, vl.z); split_fast(v3, vh.w, vl.w);
                        const unsigned off = rbase + ((unsigned)k4 ^ (unsigned)(st & 7)) * 16u;
                        *reinterpret_cast<float4 *>(hi + off) = vh;
                        *reinterpret_cast<float4 *>(lo + off) = vl;
                    }
                }
                GS_T(8, sta == 0);
                fence_proxy_async();               // generic-proxy writes -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(&ring_full[s]);        // one arrival per warp: 128 arrivals on one barrier word serialise
                GS_T(9, sta == 0);
                if (st == 0) GS_TR(grp, it, c, 1);
            }
            if (use_tma) {
                fence_proxy_async();               // the in-place gamma scaling is ordered before the next bulk copy into this buffer
                __syncwarp();
                if (lane == 0) mbar_arrive(&raw_empty[buf]);
            }
            GS_T(10, sta == 0);
        }
    } else {
        // ---------------- epilogue: thread = TMEM lane = row m of Q, the two warp groups split the columns.  Stored transposed
        // (Q is symmetric): a warp writes 32 consecutive floats of row n per instruction.
        const int m = 32 * (warp & 3) + lane;
        const int half = (BN / 2 + 7) / 8 * 8;
        const int cbeg = (warp >> 2) * half, cend = min(min(BN, cbeg + half), (N + 7) / 8 * 8);
        const int ldq = p.LDQ;
        long long it = 0;
        for (long long b = b_first; b < p.B; b += b_step, ++it) {
            const unsigned tb = (unsigned)(it & 1);
            GS_T0
            mbar_wait(&acc_full[tb], (unsigned)((it >> 1) & 1));
            tc::fence_after_sync();
            GS_T(11, tid == 0);
            if (tid == 0) GS_TR(3, it, 0, 0);
            if (32 * (warp & 3) < N) {
                const uint32_t lane_addr = tmem + tb * 2u * (uint32_t)BN + ((uint32_t)(32 * (warp & 3)) << 16);
                float *qcol = p.Q + (b * (long long)N + cbeg) * ldq + m;
#pragma unroll 1
                for (int c0 = cbeg; c0 < cend; c0 += 16, qcol += 16 * ldq) {
                    float h[16], x[16];
                    const bool two = c0 + 8 < cend;
                    tmem_ld8_nowait(lane_addr + (uint32_t)c0, *reinterpret_cast<float(*)[8]>(&h[0]));
                    tmem_ld8_nowait(lane_addr + (uint32_t)(BN + c0), *reinterpret_cast<float(*)[8]>(&x[0]));
                    if (two) {
                        tmem_ld8_nowait(lane_addr + (uint32_t)(c0 + 8), *reinterpret_cast<float(*)[8]>(&h[8]));
                        tmem_ld8_nowait(lane_addr + (uint32_t)(BN + c0 + 8), *reinterpret_cast<float(*)[8]>(&x[8]));
                    }
                    tmem_ld_wait();
                    if (m < N) {
                        if (two && c0 + 16 <= N) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) qcol[j * ldq] = gsq * (h[j] + x[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (c0 + j < N && (j < 8 || two)) qcol[j * ldq] = gsq * (h[j] + x[j]);
                        }
                    }
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[tb]);
            GS_T(12, tid == 0);
            if (tid == 0) GS_TR(3, it, 0, 1);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kGsMmaWarp) tc::tmem_dealloc(tmem, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------------
// (2) one warp per portfolio
// ------------------------------------------------------------------------------------------------------------------
constexpr int kQwFree = 32;            // largest reduced system (one row per lane)
constexpr int kQwAssets = 4;           // assets per lane: n_assets <= 128
constexpr int kQwTrim = 16;            // largest overflow of the frame that is trimmed instead of handed to the dense kernel
constexpr int kQwTrims = 4;            // ... and how often per portfolio
constexpr int kQwOver = 256;           // portfolios per call that may take the list-mode CTA kernel instead of the dense one

struct QwParams {
    const float *rets, *Q, *alpha;
    float u;
    long long B;
    int N, LDQ;
    float *w;
    int8_t *state;
    int32_t *status;
    int *fail_count, *fail_list, *work_count, *work_list, *bad_count;
    const int *skip;        // probe flag: hand every portfolio to the dense kernel
    int *over_count, *over_list;   // portfolios whose sets overflow the frame: the one-CTA-per-portfolio kernel in list mode
};

// shared memory of one warp: mbarrier | multipliers L (32 x 33) | rets | reduced-row -> asset, -> slot | 32 row slots of Q
// (multipliers packed as the strict lower triangle: 496 floats; row slots hold n_assets rounded up to 4 floats, not the scratch's
// row stride -- together 14 instead of 12 warps per SM at n_assets = 100)
constexpr int kQwTri = kQwFree * (kQwFree - 1) / 2;
__host__ __device__ constexpr int qw_tri_off(int k) { return (kQwFree - 1) * k - k * (k - 1) / 2; }      // row k holds l(i, k), i = k+1 .. 31
__host__ __device__ inline size_t qw_smem_bytes(int RS) { return 16 + (size_t)kQwTri * 4 + (size_t)RS * 4 + 256 + (size_t)kQwFree * RS * 4; }

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
    if (lane > K) Lt[qw_tri_off(K) + lane - K - 1] = l;
#pragma unroll
    for (int j = K + 1; j < kQwFree; ++j) {
        const float rk = __shfl_sync(kFullMask, v[j], K);
        v[j] = fmaf(-l, rk, v[j]);
    }
    const float y1k = __shfl_sync(kFullMask, y1, K), y2k = __shfl_sync(kFullMask, y2, K);
    y1 = fmaf(-l, y1k, y1); y2 = fmaf(-l, y2k, y2);
}

__global__ void __launch_bounds__(32, 14) markowitz_qwarp_kernel(QwParams p) {
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
    if (probe_says_dense(p.skip)) {         // diversified batch (probe kernel): the dense kernel's work list
        if (lane == 0) { const int sl = atomicAdd(p.work_count, 1); p.work_list[sl] = (int)b; }
        return;
    }
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(qw_smem);
    const int RS = (N + 3) & ~3;                         // row slot stride (floats): 16-byte multiples for the bulk copies
    float *Lt = reinterpret_cast<float *>(qw_smem + 16);
    float *rs = Lt + kQwTri;
    int *idx = reinterpret_cast<int *>(rs + RS);
    int *sls = idx + kQwFree;
    float *rows = reinterpret_cast<float *>(sls + kQwFree);
    const unsigned row_bytes = (unsigned)RS * 4u;
    if (lane == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    const float *Qb = p.Q + b * (long long)N * LDQ;      // Q = A'A without the |alpha| I term, added where the diagonal is used
    const float aabs = fabsf(p.alpha[b]);
    const unsigned lt = (1u << lane) - 1u;
    float rr[kQwAssets], i2q[kQwAssets], wv[kQwAssets];
    int st[kQwAssets], slot[kQwAssets];
    bool in[kQwAssets];
    // ---- warm start: safeguarded Newton on the multiplier of the separable problem with diag(Q) (diag_init_warp)
    float dmax = 0.f, nu0 = 0.f;
    {
        float lo = Num<float>::inf(), hi = -Num<float>::inf(), sa = 0.f, sb = 0.f;
#pragma unroll
        for (int m = 0; m < kQwAssets; ++m) {
            const int i = lane + 32 * m;
            in[m] = i < N;
            rr[m] = 0.f; i2q[m] = 0.f; wv[m] = 0.f; st[m] = 2; slot[m] = -1;
            if (in[m]) {
                float q = Qb[i * LDQ + i] + aabs;
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
        nu0 = nu;
    }
    const float dmin = 4.f * Num<float>::eps() * dmax + Num<float>::tiny();
    const float tolw = 8.f * Num<float>::eps();
    int lifted = 0, converged = 0, bail = 0, trimmed = 0;
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
        // the same in every lane -- said in a way the compiler can see (a broadcast from lane 0 is warp-uniform by construction):
        // without it every shuffle below is compiled a second time in its slow divergent-safe form and the kernel no longer
        // fits the instruction cache
        nF = __shfl_sync(kFullMask, nF, 0); nU = __shfl_sync(kFullMask, nU, 0);
#pragma unroll
        for (int m = 0; m < kQwAssets; ++m) mu[m] = __shfl_sync(kFullMask, mu[m], 0);
        if (nF + nU > kQwFree) {
            // A slightly oversized free set (the warm start overshoots: 22.6 free assets on average against 19.3 at the optimum
            // of the headline workload) is cut down to the frame instead of handing the portfolio to the dense kernel, whose
            // latency for a handful of portfolios is as long as this whole kernel: the free assets with the smallest weights go
            // to the lower bound, and the iteration frees them again if their gradient says so.  The converged point is the same
            // KKT point; only the path differs.  Anything larger, or a fifth overflow, is the dense kernel's (one hand-over costs the whole
            // batch ~70 us of dense-kernel latency: on 8 GPUs one rank in eight paid it, 0.86 weak-scaling efficiency).
            const int drop = nF + nU - kQwFree;
            if (drop > kQwTrim || trimmed >= kQwTrims) { bail = 1; break; }
            ++trimmed;
            for (int d = 0; d < drop; ++d) {
                float best = Num<float>::inf();
                int bi = 0x7fffffff;
#pragma unroll
                for (int m = 0; m < kQwAssets; ++m) {
                    const float key = iter == 0 ? (rr[m] - nu0) * i2q[m] : wv[m];
                    if (st[m] == 0 && key < best) { best = key; bi = lane + 32 * m; }
                }
                warp_argmin(best, bi);
#pragma unroll
                for (int m = 0; m < kQwAssets; ++m)
                    if (lane + 32 * m == bi) st[m] = -1;
            }
            --iter;                                         // lists again with the trimmed set (not an active-set iteration)
            continue;
        }
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
            nnew = __shfl_sync(kFullMask, nnew, 0);
            __syncwarp();                                  // every read of the slots that are handed on has been issued
            if (nnew > 0) {
                if (lane == 0) mbar_expect_tx(mbar, (unsigned)nnew * row_bytes);
                __syncwarp();
#pragma unroll
                for (int m = 0; m < kQwAssets; ++m)
                    if ((needm[m] >> lane) & 1u) tma_bulk_g2s(rows + slot[m] * RS, Qb + (lane + 32 * m) * LDQ, row_bytes, mbar);
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
                        v[j] = (live && j >= k0) ? rows[sj * RS + a] : 0.f;
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
                        if (live) acc += rows[sa * RS + 32 * m + bit];
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
            const float *Lrow = Lt + qw_tri_off(lane) - lane - 1;      // l(i, lane) at Lrow[i]
            for (int i = kQwFree - 1; i > k0; --i) {
                const float xi = __shfl_sync(kFullMask, x, i);
                if (live && lane < i) x = fmaf(-Lrow[i], xi, x);
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
            const float *row = rows + sj * RS + lane;
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
                    const float *row = rows + sc * RS + lane;
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
        // a set that does not fit the 32-row frame: the one-CTA-per-portfolio kernel (<= 48 candidates) takes it in list mode
        // -- a third of the dense kernel's latency, which the whole batch waits for; beyond kQwOver of them, the dense kernel
        if (lane == 0) {
            const int ov = atomicAdd(p.over_count, 1);
            if (ov < kQwOver) p.over_list[ov] = (int)b;
            else { const int sl = atomicAdd(p.work_count, 1); p.work_list[sl] = (int)b; }
        }
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
// (0) probe: is this batch in the diversified regime?
// ------------------------------------------------------------------------------------------------------------------
// When most assets are free (small expected returns: what BachelierNet / ThorpNet produce) every portfolio outgrows the 32-row
// frame and is the dense kernel's anyway; the Gram kernel would then cost 110 us per 4096 portfolios for nothing.  The probe runs
// the warm start (column norms + Newton on the multiplier, as markowitz_sparse.cuh does) on up to kProbe portfolios spread over
// the batch and raises a flag when three quarters of them overflow the frame; the two kernels below then return at once and the
// whole batch goes to the dense kernel.  The decision is a function of the inputs of THIS call only (repeatable, capturable).
constexpr int kProbe = 32;
constexpr int kProbeMargin = 4;            // warm-start sets up to 32 + 4 assets count as "fits the warp solver's frame"
constexpr int kProbeThreads = 1024;       // n_assets <= 128: at least 8 row slices per column, at most 16 rows per thread

__global__ void __launch_bounds__(kProbeThreads) markowitz_qprobe_kernel(const float *C, const float *rets, const float *gamma, const float *alpha,
                                                                         float u, long long B, long long Bmod, long long b0, int N, int nprobe,
                                                                         int *word, int *over_count) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *over_count = 0;      // (first kernel of the call: clears the overflow list)
    __shared__ float part[kProbeThreads], qd[128 + 32], r[128 + 32];
    __shared__ int8_t st[128 + 32];
    const int tid = threadIdx.x;
    const long long b = (long long)blockIdx.x * B / nprobe;        // probes spread evenly over the batch
    // column norms of the gamma-tiled covmat_sqrt straight from global memory: thread = (column i, row slice h), coalesced rows
    const int nh = kProbeThreads / N, i = tid % N, h = tid / N;
    float acc = 0.f;
    if (h < nh) {
        const float *col = C + b * (long long)N * N + i;
        const unsigned uB = (unsigned)Bmod;
        unsigned gi = (unsigned)(((b0 + b) * (long long)N * N + (long long)h * N + i) % Bmod);
        const unsigned gstep = (unsigned)(((long long)nh * N) % Bmod);
        float cv[16], gv[16];               // every load of this thread in flight at once: the probe is latency, not bandwidth
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int k = h + q * nh;
            cv[q] = k < N ? __ldg(col + (long long)k * N) : 0.f;
            gv[q] = k < N ? __ldg(gamma + gi) : 0.f;
            gi += gstep; if (gi >= uB) gi -= uB;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) { const float v = cv[q] * gv[q]; acc = fmaf(v, v, acc); }
    }
    part[tid] = acc;
    if (tid < N) r[tid] = rets[b * N + tid];
    __syncthreads();
    if (tid < N) {
        float q = fabsf(alpha[b]);
        for (int hh = 0; hh < nh; ++hh) q += part[hh * N + tid];
        qd[tid] = q;
    }
    __syncthreads();
    if (tid < 32) {
        float dmax;
        diag_init_warp<float, kQwAssets>(qd, 1, r, st, N, u, &dmax);
        __syncwarp();
        int held = 0;
        for (int j = tid; j < N; j += 32) held += st[j] >= 0;       // free or capped: the assets that need a row of the frame
        held = warp_sum(held);
        if (tid == 0) {
            // a warm start a few assets over the frame usually shrinks into it (22.6 -> 19.3 free on the headline workload: that
            // is what the trimming is for); one well beyond it stays beyond it
            const int mine = (1 << 16) | (held > kQwFree + kProbeMargin ? 1 : 0);
            const int seen = atomicAdd(word, mine) + mine;
            if ((seen >> 16) == nprobe && 4 * (seen & 0xffff) >= 3 * nprobe) atomicOr(word, (int)0x80000000);
        }
    }
}

#ifdef DDW_GS_TIMING
}  // namespace ddw
extern "C" int ddw_debug_gs_phases(unsigned long long *out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, ddw::ddw_gs_cycles, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(ddw::ddw_gs_cycles, z, sizeof(z)); }
    return 0;
}
extern "C" int ddw_debug_gs_trace(long long *out4096) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out4096, ddw::ddw_gs_trace, sizeof(long long) * 4096);
    return 0;
}
namespace ddw {
#endif

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
bool fwd_qwarp_eligible(int N, int tsize) {
    static const bool off = getenv("DDW_NO_QWARP") != nullptr;      // A/B switch: the one-CTA-per-portfolio kernel instead
    return !off && tsize == 4 && N >= 8 && N <= 128;
}
int fwd_qwarp_ldq(int N) { return (N + 7) / 8 * 8; }     // rows of the scratch start on 32-byte sectors
// Q (B x N x LDQ) followed by the overflow list (count + kQwOver entries)
size_t fwd_qwarp_scratch_bytes(long long B, int N) { return (size_t)B * N * fwd_qwarp_ldq(N) * sizeof(float) + 2048; }

template <typename T> int launch_fwd_sparse(const MkFwdParams<T> &p, cudaStream_t st);      // markowitz_sparse.cuh
template <typename T> int launch_fwd_sparse_list(const MkFwdParams<T> &p, const int *in_list, const int *in_count, int max_entries, cudaStream_t st);

// Below this batch size the one-CTA-per-portfolio kernel is at least as fast: one wave of it (592 portfolios) takes about as long as
// the probe + one portfolio's trip through the Gram pipeline + one warp's solve (measured at B = 64 .. 1024, tools/time_configs.py).
static long long qwarp_min_batch() {
    const char *e = getenv("DDW_QWARP_MIN_BATCH");
    return e ? atoll(e) : 1024;
}

int launch_fwd_qwarp(const MkFwdParams<float> &p, float *Q, cudaStream_t st) {
    if (p.B < qwarp_min_batch()) return launch_fwd_sparse<float>(p, st);
    static DeviceOnce configured;
    const int dev = current_device();
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const GsLayout l = gs_layout(p.N, device_smem_limit());
    if (configured.needs(dev)) {
        if (cudaFuncSetAttribute(markowitz_gram_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, device_smem_limit()) != cudaSuccess ||
            cudaFuncSetAttribute(markowitz_qwarp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qw_smem_bytes(128)) != cudaSuccess)
            return check_launch("cudaFuncSetAttribute(markowitz_gram_small_kernel / markowitz_qwarp_kernel)");
        configured.mark(dev);
    }
    int *probe = p.fail_count + 3;          // fourth counter of the workspace header, cleared with the other three by the caller
    int *over = reinterpret_cast<int *>(Q + (size_t)p.B * p.N * fwd_qwarp_ldq(p.N));      // [0] count, [1 .. kQwOver] portfolios
    {
        const int nprobe = (int)(p.B < kProbe ? p.B : kProbe);
        markowitz_qprobe_kernel<<<nprobe, kProbeThreads, 0, st>>>(p.C, p.rets, p.gamma, p.alpha, p.u, p.B, p.Bmod, p.b0, p.N, nprobe, probe, over);
        int rcp = check_launch("markowitz_qprobe_kernel");
        if (rcp) return rcp;
    }
    GsParams g;
    g.skip = probe;
    g.C = p.C; g.gamma = p.gamma; g.alpha = p.alpha; g.Q = Q; g.B = p.B; g.Bmod = p.Bmod; g.b0 = p.b0;
    g.N = p.N; g.LDQ = fwd_qwarp_ldq(p.N); g.BN = l.BN; g.nraw = l.nraw;
    g.raw_bytes = l.raw_bytes; g.tile_bytes = l.tile_bytes; g.ring_bytes = l.ring_bytes; g.tmem_cols = l.tmem_cols;
    const long long grid = nsm > 0 ? nsm : 148;          // one persistent CTA per SM (it owns the SM's tensor memory)
    markowitz_gram_small_kernel<<<(unsigned)(p.B < grid ? p.B : grid), kGsThreads, l.total, st>>>(g);
    int rc = check_launch("markowitz_gram_small_kernel");
    if (rc) return rc;
    QwParams q;
    q.skip = probe;
    q.rets = p.rets; q.Q = Q; q.alpha = p.alpha; q.u = p.u; q.B = p.B; q.N = p.N; q.LDQ = g.LDQ; q.w = p.w; q.state = p.state; q.status = p.status;
    q.fail_count = p.fail_count; q.fail_list = p.fail_list; q.work_count = p.work_count; q.work_list = p.work_list;
    q.bad_count = p.bad_count;
    q.over_count = over; q.over_list = over + 1;
    markowitz_qwarp_kernel<<<(unsigned)p.B, 32, qw_smem_bytes((p.N + 3) & ~3), st>>>(q);
    rc = check_launch("markowitz_qwarp_kernel");
    if (rc) return rc;
    return launch_fwd_sparse_list<float>(p, over + 1, over, kQwOver, st);
}

}  // namespace ddw
