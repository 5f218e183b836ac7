// tc_gemm.cuh -- 5th-generation tensor-core (tcgen05 + TMEM) building blocks and a tile GEMM in 3xTF32.
//
// Used for the only GEMM-shaped work on deepdow's allocation path: the X'X contraction of
// CovarianceMatrix.compute_covariance (reference deepdow/layers/misc.py:147) at large n_assets, the
// rotation updates of the blocked Jacobi eigensolver behind compute_sqrt (misc.py:183-201) and the
// v sqrt(s) v' reconstruction (misc.py:201).
//
// One CTA computes one 128 x BN output tile  C(m, n) = sum_k A(m, k) * B(n, k)  ("NT" form: both operands are
// addressed as (row, k)).  Operands are read from global memory by all threads through a loader functor, split
// into a TF32 "hi" part and the exact FP32 remainder "lo" (a == hi + lo; the tensor core reads the top 10 mantissa bits of
// lo, so the product is good to 2^-21), and written to shared memory in the tensor core's canonical K-major no-swizzle
// layout (8 x 16-byte core matrices).  One elected thread issues tcgen05.mma.kind::tf32 three times per k-step
// (lo*hi + hi*lo + hi*hi) into FP32 accumulators in TMEM; they are read back with tcgen05.ld (one TMEM lane = one output
// row per thread) and handed to an epilogue functor column by column, so that "transposed" stores C[n][m] are coalesced
// across the warp.  The k loop is double buffered: the MMAs of chunk c run while the threads stage chunk c + 1.
//
// Accumulation: the tensor core adds into TMEM with round-toward-zero, once per instruction (measured on a B200 with
// tools/tc_probe.cu: one accumulator gave 1.2e-5 absolute error at K = 250 on sums of magnitude 6, exactly what a
// truncating model predicts, tools/models/block_jacobi_model.py).  The tile therefore keeps FOUR accumulators: the
// hi*hi products of k-step s go to accumulator s % 3, both cross terms to the fourth, and the epilogue adds the four in
// FP32 round-to-nearest.  That brings the error below a plain FP32 FMA loop's (model: 1.7e-6 -> measured, see profiles/).
#pragma once
#include "common.cuh"

namespace ddw {
namespace tc {

constexpr int kTileM = 128;          // rows of A per tile = TMEM lanes
constexpr int kChunkK = 32;          // k values staged per pipeline stage
constexpr int kThreads = 256;
constexpr int kLbo = 144;            // bytes between K-adjacent core matrices (128 B of data + 16 B of bank skew)
constexpr int kSbo = (kChunkK / 4) * kLbo;   // bytes between MN-adjacent 8-row groups (pipelined tile, chunks of 32 k)
constexpr int kAccs = 4;             // TMEM accumulators per tile: 3 for hi*hi (round robin over k-steps) + 1 for the cross terms

// ---- raw tcgen05 / TMEM primitives (PTX ISA; SASS: UTCHMMA-family, LDTM, UTCBAR) ------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_smem, uint32_t ncols) {   // one whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {        // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// instruction descriptor of tcgen05.mma.kind::tf32: D = F32, A = B = TF32, both K-major, M x N tile
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// shared-memory matrix descriptor, K-major, no swizzle: 8-row x 16-byte core matrices, `lbo` bytes between core
// matrices adjacent in K, `sbo` bytes between core matrices adjacent in M/N
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on `bar` once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive FP32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

// 16 consecutive FP32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// a == hi + lo exactly: hi = a rounded to TF32 (10-bit mantissa), lo = the FP32 remainder (13 significant bits at most)
__device__ __forceinline__ void split_tf32(float a, float &hi, float &lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(a));
    hi = __uint_as_float(h);
    lo = a - hi;
}

// bytes of dynamic shared memory of tile_gemm<BN>: two stages of {A hi, A lo, B hi, B lo} + 1 KB of alignment slack
template <int BN> constexpr size_t tile_gemm_smem() {
    return (size_t)2 * (2 * (kTileM / 8) * kSbo + 2 * (BN / 8) * kSbo) + 1024;
}

struct TileShared {            // static shared state of one tile GEMM
    unsigned long long done[2], fin;
    uint32_t tmem_base;
};

// Stage `ROWS` x KT elements of one operand (rows r0.., k values k0..) as hi/lo tiles; element (r, k) lands at byte
// (r / 8) * (KT / 4) * kLbo + (k / 4) * kLbo + (r % 8) * 16 + (k % 4) * 4 of each tile.
// Loader: void ld4(int r, int k, float (&v)[4]) returns elements (r, k .. k+3), zero outside the operand, and
// static constexpr bool kRowsFast: true when consecutive r are adjacent in memory (lanes then walk r), false when
// consecutive k are (lanes walk k in groups of four).
template <int ROWS, int KT, typename Loader, int kPer = 4>
__device__ __forceinline__ void stage_operand(const Loader &ldr, int r0, int k0, unsigned char *hi, unsigned char *lo) {
    constexpr int kItems = ROWS * (KT / 4);               // kPer = items (4 elements each) in flight per thread and pass
    constexpr int kSboT = (KT / 4) * kLbo;
    static_assert(kItems % (kThreads * kPer) == 0 || kItems % kThreads == 0, "tile does not divide over the CTA");
    constexpr int kPasses = kItems / kThreads;
#pragma unroll
    for (int p0 = 0; p0 < kPasses; p0 += kPer) {
        float v[kPer][4];
#pragma unroll
        for (int it = 0; it < kPer; ++it) {
            if (p0 + it < kPasses) {
                const int id = (p0 + it) * kThreads + threadIdx.x;
                int r, k4;
                if (Loader::kRowsFast) { r = id % ROWS; k4 = id / ROWS; }
                else { k4 = id & (KT / 4 - 1); r = id / (KT / 4); }
                ldr.ld4(r0 + r, k0 + 4 * k4, v[it]);
            }
        }
#pragma unroll
        for (int it = 0; it < kPer; ++it) {
            if (p0 + it < kPasses) {
                const int id = (p0 + it) * kThreads + threadIdx.x;
                int r, k4;
                if (Loader::kRowsFast) { r = id % ROWS; k4 = id / ROWS; }
                else { k4 = id & (KT / 4 - 1); r = id / (KT / 4); }
                float4 h, l;
                split_tf32(v[it][0], h.x, l.x); split_tf32(v[it][1], h.y, l.y);
                split_tf32(v[it][2], h.z, l.z); split_tf32(v[it][3], h.w, l.w);
                const int off = (r >> 3) * kSboT + k4 * kLbo + (r & 7) * 16;
                *reinterpret_cast<float4 *>(hi + off) = h;
                *reinterpret_cast<float4 *>(lo + off) = l;
            }
        }
    }
}

// the 3 x TF32 MMAs of `ksteps` k-steps (8 k values each) of one staged tile pair; `step0` = index of the first k-step
// in the whole contraction (selects the accumulators and the very first overwrite)
template <int BN>
__device__ __forceinline__ void issue_mma3(uint32_t tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                           int ksteps, int step0, uint32_t lbo, uint32_t sbo) {
    constexpr uint32_t idesc = idesc_tf32(kTileM, BN);
    for (int kk = 0; kk < ksteps; ++kk) {
        const uint32_t o = kk * 2 * lbo;       // one MMA consumes 8 k values = 2 core matrices along K
        const uint64_t dah = smem_desc(a_hi + o, lbo, sbo), dal = smem_desc(a_lo + o, lbo, sbo);
        const uint64_t dbh = smem_desc(b_hi + o, lbo, sbo), dbl = smem_desc(b_lo + o, lbo, sbo);
        const int s = step0 + kk;
        mma_tf32(tmem + 3 * BN, dal, dbh, idesc, s != 0);
        mma_tf32(tmem + 3 * BN, dah, dbl, idesc, 1u);
        mma_tf32(tmem + (s % 3) * BN, dah, dbh, idesc, s >= 3);
    }
}

// sum of the four accumulators: 32 consecutive columns (from c0) of this thread's TMEM lane
template <int BN>
__device__ __forceinline__ void load_acc32(uint32_t tmem, int warp, int c0, float (&v)[32]) {
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0;
    tmem_ld32(lane_addr, v);
#pragma unroll
    for (int a = 1; a < kAccs; ++a) {
        float t[32];
        tmem_ld32(lane_addr + a * BN, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += t[j];
    }
}

// One 128 x BN output tile.  Every thread of the CTA (kThreads) must call it.  `lbo`/`sbo` are passed through so
// that the probe (tools/tc_probe.cu) can exercise the descriptor convention.
// Epilogue: void operator()(int m, int n, float v) for 0 <= m - m0 < 128, 0 <= n - n0 < BN; for a fixed n the 32 lanes of a
// warp hold 32 consecutive m.
template <int BN, typename LoaderA, typename LoaderB, typename Epilogue>
__device__ __forceinline__ void tile_gemm(const LoaderA &la, const LoaderB &lb, int m0, int n0, int K, Epilogue epi,
                                          unsigned char *smem_dyn, TileShared &sh, uint32_t lbo = kLbo,
                                          uint32_t sbo = kSbo) {
    static_assert(BN % 32 == 0 && BN >= 32 && kAccs * BN <= 512, "BN");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kABytes = (kTileM / 8) * kSbo, kBBytes = (BN / 8) * kSbo;
    constexpr int kStage = 2 * kABytes + 2 * kBBytes;
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t kCols = kAccs * BN <= 128 ? 128 : (kAccs * BN <= 256 ? 256 : 512);
    if (tid == 0) { mbar_init(&sh.done[0], 1); mbar_init(&sh.done[1], 1); mbar_init(&sh.fin, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&sh.tmem_base, kCols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = sh.tmem_base;
    const int nchunks = (K + kChunkK - 1) / kChunkK;
    for (int c = 0; c < nchunks; ++c) {
        const int s = c & 1;
        unsigned char *st = base + s * kStage;
        // the MMAs that read this stage two chunks ago must have drained
        if (c >= 2) mbar_wait(&sh.done[s], ((c >> 1) + 1) & 1);
        stage_operand<kTileM, kChunkK>(la, m0, c * kChunkK, st, st + kABytes);
        stage_operand<BN, kChunkK>(lb, n0, c * kChunkK, st + 2 * kABytes, st + 2 * kABytes + kBBytes);
        fence_proxy_async();           // generic-proxy writes -> visible to the tensor core's async proxy
        __syncthreads();
        if (tid == 0) {
            fence_after_sync();
            const uint32_t a_hi = smem_u32(st), a_lo = a_hi + kABytes, b_hi = a_hi + 2 * kABytes, b_lo = b_hi + kBBytes;
            issue_mma3<BN>(tmem, a_hi, a_lo, b_hi, b_lo, kChunkK / 8, c * (kChunkK / 8), lbo, sbo);
            mma_commit(&sh.done[s]);
            if (c == nchunks - 1) mma_commit(&sh.fin);
        }
    }
    mbar_wait(&sh.fin, 0);
    fence_after_sync();
    // epilogue: warps 0..3 and 4..7 split the columns, each thread owns TMEM lane 32 * (warp % 4) + lane
    {
        const int m = m0 + 32 * (warp & 3) + lane;
        constexpr int kHalf = BN / 2 >= 32 ? BN / 2 : BN;
        const int cbeg = (BN / 2 >= 32) ? (warp >> 2) * kHalf : 0;
        if (BN / 2 >= 32 || warp < 4) {
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + kHalf; c0 += 32) {
                float v[32];
                load_acc32<BN>(tmem, warp, c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) epi(m, n0 + c0 + j, v[j]);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kCols);
}

// ---- loaders ---------------------------------------------------------------------------------------------------
// element (r, k) = p[r * ld + k], zero outside [0, R) x [0, K); 16-byte loads when the row segment is aligned and inside
struct LoadKContig {
    static constexpr bool kRowsFast = false;
    const float *p; long long ld; int R, K;
    __device__ __forceinline__ void ld4(int r, int k, float (&v)[4]) const {
        if (r < R && k + 3 < K && (((reinterpret_cast<uintptr_t>(p) >> 2) + r * ld + k) & 3) == 0) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(p + r * ld + k));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = (r < R && k + q < K) ? __ldg(p + r * ld + k + q) : 0.f;
        }
    }
};
// element (r, k) = (p[k * ld + r] - sub[r]) * scale[k]  (sub / scale optional), zero outside
struct LoadRowsContig {
    static constexpr bool kRowsFast = true;
    const float *p; long long ld; int R, K;
    const float *sub, *scale;
    __device__ __forceinline__ void ld4(int r, int k, float (&v)[4]) const {
        const float s = (sub && r < R) ? sub[r] : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float x = (r < R && k + q < K) ? __ldg(p + (long long)(k + q) * ld + r) - s : 0.f;
            if (scale && k + q < K) x *= scale[k + q];
            v[q] = x;
        }
    }
};

}  // namespace tc
}  // namespace ddw
