// markowitz_sparse.cuh -- forward solver for portfolios with a small support
// (compiled in markowitz_fwd_sparse_f32.cu / markowitz_fwd_sparse_f64.cu; problem statement in markowitz_impl.cuh)
//
// The dense kernel forms the whole Gram Q = A'A (N^3/2 flops) although an active-set iteration only
// ever needs (i) the block of Q on the assets that are or were free and (ii) gradients 2 A'(A w).
// Here A itself (gamma-tiled covmat_sqrt, 4 N^2 bytes = one contiguous block of HBM) is brought into
// shared memory by ONE bulk TMA copy per portfolio, and Q entries are computed only for assets when
// they first become free ("candidates").  Work per portfolio drops from ~N^3/2 to
// ~|cand|^2 N/2 + iterations * 2 N^2 flops.  Portfolios whose candidate set outgrows kCandMax are
// appended to a work list that the dense kernel then processes.
//
// Per active-set iteration (one CTA of 128 threads per portfolio):
//   lists      every warp compacts its 32 assets of the free / capped sets and hands out candidate slots
//   Gram       4x4 register tiles of a_s . a_c for the new candidates, rows split over 8 lanes per tile
//   KKT solve  all four warps factor the augmented triangle with every element in a register (one barrier per
//              column); warp 0 then does the multiplier and the back substitution
//   gradient   t = A w on the support, then g = 2 A't as a column pass with 16-byte loads, rows split over
//              the thread groups and combined through shared memory
#pragma once
#include <stdlib.h>
#include "markowitz_common.cuh"
#include "support_kernels.cuh"

namespace ddw {

constexpr int kCandMax = 48;
constexpr int kCandLd = (kCandMax + 3) | 1;
constexpr int kCandSlots = (kCandMax + 31) / 32;   // vector elements per lane in the one-warp back substitution

template <typename T> struct SpSmem {
    unsigned long long *mbar;
    T *As, *Qc, *part, *r, *wv, *tv, *g, *qd, *dinv, *wF, *red;   // g aliases dinv (only used when no asset is free)
    int *idx, *idxU, *slot, *sF, *cand, *misc, *cnt;
    int8_t *st;
};

constexpr int sp_round_up(int v, int m) { return (v + m - 1) / m * m; }
constexpr size_t sp_smem_bytes(int N, int tsize) {
    return 16 + ((size_t)sp_round_up(N * N, 4) + sp_round_up(kCandMax * kCandLd, 4) + kPartElems + 6 * (size_t)sp_round_up(N + 2, 32) + 40) * tsize +
           ((size_t)3 * sp_round_up(N + 2, 32) + 64 + 64 + 8) * sizeof(int) + sp_round_up(N + 2, 32);
}
// the headline shape (n_assets = 100, float) lives on 4 resident CTAs per SM: 228 KB of shared memory per SM, 1 KB
// reserved per CTA.  64 bytes more once cost 40 us per 4096 portfolios (measured) -- keep it a compile-time fact.
static_assert(4 * (sp_smem_bytes(100, 4) + 1024) <= 228 * 1024, "sparse forward kernel: n_assets=100 must fit 4 CTAs per SM");

#ifdef DDW_MK_DEFINE_ELIGIBLE
// sparse-support kernel first whenever A (4 N^2 bytes) fits shared memory with room for >= 2 CTAs per SM
bool fwd_sparse_eligible(int N, int tsize, int smem_limit) {
    return N <= 128 && N >= 8 && sp_smem_bytes(N, tsize) <= (size_t)smem_limit / 2;
}
#else
bool fwd_sparse_eligible(int N, int tsize, int smem_limit);
#endif

template <typename T> __device__ __forceinline__ SpSmem<T> sp_carve(unsigned char *raw, int N) {
    SpSmem<T> s;
    const int NP = round_up(N + 2, 32);
    s.mbar = reinterpret_cast<unsigned long long *>(raw);
    T *p = reinterpret_cast<T *>(raw + 16);
    s.As = p; p += round_up(N * N, 4);
    s.Qc = p; p += round_up(kCandMax * kCandLd, 4);
    s.part = p; p += kPartElems;
    s.r = p; p += NP;
    s.wv = p; p += NP;
    s.tv = p; p += NP;
    s.qd = p; p += NP;
    s.dinv = p; s.g = p; p += NP;
    s.wF = p; p += NP;
    s.red = p; p += 40;
    int *q = reinterpret_cast<int *>(p);
    s.idx = q; q += NP;
    s.idxU = q; q += NP;
    s.slot = q; q += NP;
    s.cand = q; q += 64;
    s.sF = q; q += 64;
    s.misc = q; q += 8;
    s.cnt = reinterpret_cast<int *>(s.red);   // 12 ints over red[0..], which the sparse kernel uses only from index 32 up:
                                              // 64 more bytes would drop the SM from 4 resident CTAs to 3 (4 x (57.3 KB + 1 KB) = 228 KB)
    s.st = reinterpret_cast<int8_t *>(q);
    return s;
}

template <typename T, int kThreads, int NREG>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? 4 : 2) markowitz_fwd_sparse_kernel(MkFwdParams<T> p, const int *in_list,
                                                                                               const int *in_count) {
    static_assert(kThreads == 128, "the list phase maps one thread to one asset (n_assets <= 128)");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // list mode (in_list != nullptr): CTA i takes the i-th portfolio that the warp-per-portfolio kernel could not hold in its
    // 32-row frame (markowitz_qwarp.cu); direct mode: one CTA per portfolio of the batch
    if (in_list && (int)blockIdx.x >= *in_count) return;
    const long long b = in_list ? (long long)in_list[blockIdx.x] : (long long)blockIdx.x;
    const int N = p.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int LDC = kCandLd;
    const T u = p.u;
    T *w_out = p.w + b * N;
    int8_t *st_out = p.state + b * N;
    if (trivial_cases<T, kThreads>(N, u, w_out, st_out, p.status + b, p.bad_count)) {
        if (p.saved_n && tid == 0) p.saved_n[b] = -1;
        return;
    }
    SpSmem<T> s = sp_carve<T>(smem_raw, N);
    T *As = s.As, *Qc = s.Qc, *Lb = s.Qc + 1;
    const T *Cb = p.C + b * (long long)N * N;
    const long long gflat0 = (p.b0 + b) * (long long)N * N;

    // ---- A -> shared memory: one bulk TMA copy (40 KB at N = 100), completion on an mbarrier
    DDW_PHASE_INIT
    DDW_PHASE_COUNT(11);
    const bool use_tma = stage_matrix_begin<T, kThreads>(As, Cb, N, s.mbar);
    for (int i = tid; i < N; i += kThreads) { s.r[i] = p.rets[b * N + i]; s.slot[i] = -1; }
    if (tid == 0) { s.misc[3] = 0; }
    const T aabs = absT(p.alpha[b]);
    stage_matrix_end<T, kThreads>(As, Cb, N, s.mbar, use_tma);
    DDW_PHASE(0);
    // ---- gamma tiling (allocate.py:268-270), in place
    apply_gamma_tiling<T, kThreads>(As, N, p.gamma, gflat0, p.Bmod);
    __syncthreads();
    // ---- diag(Q) = column norms + |alpha|
    {
        const int slices = column_pass<T, kThreads, true>(As, N, nullptr, s.part);
        __syncthreads();
        for (int i = tid; i < N; i += kThreads) s.qd[i] = column_pass_sum(s.part, N, slices, i) + aabs;
    }
    __syncthreads();
    DDW_PHASE(1);
    if (warp == 0) {
        T dmax;
        diag_init_warp<T, NREG>(s.qd, 1, s.r, s.st, N, u, &dmax);
        if (lane == 0) s.red[33] = dmax;
    }
    __syncthreads();
    DDW_PHASE(2);
    const T dmin = T(4) * Num<T>::eps() * s.red[33] + Num<T>::tiny();
    const T tolw = T(8) * Num<T>::eps();
    int lifted = 0, converged = 0, bail = 0;
#pragma unroll 1
    for (int it = 0; it < kMaxPdas; ++it) {
        // (a) lists of free / capped assets; first-time free assets get a candidate slot.  Warp w owns assets
        // 32w .. 32w+31 (n_assets <= 128 here): ballots, per-warp counts through shared memory, then each warp
        // writes its part of the lists at its prefix offset (same order as a serial scan)
        {
            const int i = tid;                       // kThreads == 128 covers every asset
            const int v = i < N ? (int)s.st[i] : 2;
            const bool need = v == 0 && s.slot[i < N ? i : 0] < 0;
            const unsigned lt = (1u << lane) - 1u;
            const unsigned mf = __ballot_sync(kFullMask, v == 0), mu = __ballot_sync(kFullMask, v == 1);
            const unsigned mn = __ballot_sync(kFullMask, need);
            if (lane == 0) { s.cnt[3 * warp] = __popc(mf); s.cnt[3 * warp + 1] = __popc(mu); s.cnt[3 * warp + 2] = __popc(mn); }
            __syncthreads();
            int of = 0, ou = 0, on = 0, nf = 0, nu = 0, nn = 0;
#pragma unroll
            for (int w2 = 0; w2 < kThreads / 32; ++w2) {
                const int cf = s.cnt[3 * w2], cu = s.cnt[3 * w2 + 1], cn = s.cnt[3 * w2 + 2];
                if (w2 < warp) { of += cf; ou += cu; on += cn; }
                nf += cf; nu += cu; nn += cn;
            }
            const int nc0 = s.misc[3];
            if (v == 0) s.idx[of + __popc(mf & lt)] = i;
            if (v == 1) s.idxU[ou + __popc(mu & lt)] = i;
            if (need) {
                const int sl = nc0 + on + __popc(mn & lt);
                if (sl < kCandMax) { s.slot[i] = sl; s.cand[sl] = i; }
            }
            __syncthreads();     // every thread has read misc[3] before it is advanced
            if (tid == 0) { s.misc[0] = nf; s.misc[1] = nu; s.misc[2] = nc0; s.misc[3] = nc0 + nn; }
        }
        for (int i = tid; i < N; i += kThreads) s.wv[i] = s.st[i] > 0 ? u : T(0);
        __syncthreads();
        DDW_PHASE(3);
        DDW_PHASE_COUNT(10);
        DDW_PHASE_ADD(12, s.misc[0]);
        if (it == 0) DDW_PHASE_ADD(14, s.misc[0]);
        const int nF = s.misc[0], nU = s.misc[1], nc0 = s.misc[2], nc1 = s.misc[3];
        if (nc1 > kCandMax) { bail = 1; break; }
        // (b) Gram entries of the new candidates against all candidates
        for (int m = tid; m < nF; m += kThreads) s.sF[m] = s.slot[s.idx[m]];   // candidate slots of the free assets
        if (nc1 > nc0) gram_columns<T, kThreads>(As, N, s.cand, s.qd, T(0), nc0, nc1, Qc, LDC, 1);
        // (c) t_U = A (u 1_U)
        if (nU > 0) {
            for (int k = tid; k < N; k += kThreads) {
                T acc = T(0);
                for (int m = 0; m < nU; ++m) acc += As[k * N + s.idxU[m]];
                s.tv[k] = u * acc;
            }
        }
        __syncthreads();
        const T c = T(1) - u * (T)nU;
        int changed = 0;
        T nu_mult = T(0);
        if (nF > 0) {
            // (d) reduced KKT system: Q_FF from the candidate block, rhs r_F - 2 A_F' t_U and 1
            auto rhs1 = [&](int cj) -> T {
                T acc = s.r[cj];
                if (nU > 0) {
                    T d0 = T(0);
                    for (int k = 0; k < N; ++k) d0 = fma(As[k * N + cj], s.tv[k], d0);
                    acc -= T(2) * d0;
                }
                return acc;
            };
            DDW_PHASE(4);
            // all four warps factor the augmented triangle [Q_FF; rhs; 1'] (every element in a register of its owning
            // thread, one block barrier per column), then warp 0 finishes: multiplier and back substitution.  A one-warp
            // factorisation was tried in three forms (rows in registers with shuffles, with shared-memory broadcasts,
            // and a loop over shared memory): a lone warp retires one instruction every 5-8 cycles on this dependent
            // code, 15k cycles per solve at 21 free assets against ~7k for the four-warp form (tools/phase_timing.py).
            auto init = [&](int i, int j) -> T {
                if (i < nF) {
                    const int sa = s.sF[i], sc = s.sF[j];
                    return Qc[max(sa, sc) * LDC + min(sa, sc)];
                }
                if (i > nF) return T(1);
                return rhs1(s.idx[j]);
            };
            if (small_system_fits<T, kThreads>(nF)) lifted |= factor_small<T, kThreads>(Lb, LDC, nF, s.dinv, dmin, init);
            else {
                for (int e = tid; e < nF * (nF + 2); e += kThreads) {
                    const int j = e / (nF + 2), i = e - j * (nF + 2);
                    if (i >= j) Lb[j * LDC + i] = init(i, j);
                }
                __syncthreads();
                lifted |= factor_ldl<T, kThreads>(Lb, LDC, nF, 2, s.dinv, dmin);
            }
            DDW_PHASE(5);
            if (warp == 0) {
                T s12 = T(0), s22 = T(0);
                for (int k = lane; k < nF; k += 32) {
                    const T y1 = Lb[k * LDC + nF], y2 = Lb[k * LDC + nF + 1], di = s.dinv[k];
                    s12 = fma(y1 * y2, di, s12);
                    s22 = fma(y2 * y2, di, s22);
                }
                s12 = warp_sum(s12); s22 = warp_sum(s22);
                const T nu = (s12 - T(2) * c) / s22;
                T zt[kCandSlots];
#pragma unroll
                for (int m = 0; m < kCandSlots; ++m) {
                    const int k = lane + 32 * m;
                    zt[m] = k < nF ? T(0.5) * (Lb[k * LDC + nF] - nu * Lb[k * LDC + nF + 1]) : T(0);
                }
                backsub_warp<T, kCandSlots>(Lb, LDC, nF, s.dinv, zt);
#pragma unroll
                for (int m = 0; m < kCandSlots; ++m) {
                    const int k = lane + 32 * m;
                    if (k < nF) { s.wv[s.idx[k]] = zt[m]; s.wF[k] = zt[m]; }
                }
                if (lane == 0) s.red[32] = nu;
            }
            __syncthreads();
            DDW_PHASE(6);
            nu_mult = s.red[32];
        }
        // (e) t = A w on the support, then the gradient 2 (A' t + |alpha| w) - r + nu
        for (int k = tid; k < N; k += kThreads) {
            T acc = nU > 0 ? s.tv[k] : T(0);
            const T *row = As + k * N;
            int m = 0;
            for (; m + 3 < nF; m += 4) {
                const int4 j4 = *reinterpret_cast<const int4 *>(s.idx + m);
                const Vec4<T> w4 = ld4(s.wF + m);
                acc = fma(row[j4.x], w4.v[0], acc); acc = fma(row[j4.y], w4.v[1], acc);
                acc = fma(row[j4.z], w4.v[2], acc); acc = fma(row[j4.w], w4.v[3], acc);
            }
            for (; m < nF; ++m) acc = fma(row[s.idx[m]], s.wF[m], acc);
            s.tv[k] = acc;
        }
        __syncthreads();
        DDW_PHASE(7);
        const int slices = column_pass<T, kThreads, false>(As, N, s.tv, s.part);
        __syncthreads();
        for (int i = tid; i < N; i += kThreads) {
            const T h = T(2) * (column_pass_sum(s.part, N, slices, i) + aabs * s.wv[i]) - s.r[i];
            if (nF > 0) {
                const T gi = h + nu_mult;
                const int8_t so = s.st[i];
                int8_t sn = so;
                if (so == 0) {
                    const T wi = s.wv[i];
                    if (wi < -tolw) sn = -1;
                    else if (wi > u + tolw) sn = 1;
                } else if (so < 0) { if (gi < T(0)) sn = 0; }
                else { if (gi > T(0)) sn = 0; }
                if (sn != so) { s.st[i] = sn; changed = 1; }
            } else s.g[i] = h;
        }
        if (nF == 0) {
            // no free asset: optimal iff max_L(-h) <= min_U(-h) and u |U| = 1 (same rule as the dense kernel)
            __syncthreads();
            if (warp == 0) {
                T hl = Num<T>::inf(), hu = Num<T>::inf();
                int il = 0x7fffffff, iu = 0x7fffffff;
                for (int i = lane; i < N; i += 32) {
                    const T h = s.g[i];
                    if (s.st[i] < 0) { if (h < hl) { hl = h; il = i; } }
                    else if (s.st[i] > 0) { if (-h < hu) { hu = -h; iu = i; } }
                }
                warp_argmin(hl, il); warp_argmin(hu, iu);
                const T ctol = T(4) * Num<T>::eps() * (T)N;
                if (lane == 0) {
                    int ch = 0;
                    if (c > ctol && il < N) { s.st[il] = 0; ch = 1; }
                    else if (c < -ctol && iu < N) { s.st[iu] = 0; ch = 1; }
                    else if (il < N && iu < N && -hl > hu) { s.st[il] = 0; s.st[iu] = 0; ch = 1; }
                    s.misc[4] = ch;
                }
            }
            __syncthreads();
            changed = s.misc[4];
        }
        const int any_change = __syncthreads_or(changed);
        DDW_PHASE(8);
        if (!any_change) { converged = 1; break; }
    }
    if (bail) {
        // support too large for this kernel: hand the portfolio to the dense kernel
        if (tid == 0) { const int sl = atomicAdd(p.work_count, 1); p.work_list[sl] = (int)b; }
        return;
    }
    int bad = 0;
    for (int i = tid; i < N; i += kThreads) {
        int8_t st = s.st[i];
        const T wi = st == 0 ? clampT(s.wv[i], T(0), u) : (st > 0 ? u : T(0));
        if (st == 0) { bad |= !(absT(s.wv[i]) < Num<T>::inf()); st = wi <= T(0) ? -1 : (wi >= u ? 1 : 0); }   // tie rule, see write_result
        w_out[i] = wi;
        st_out[i] = st;
    }
    const int nonfinite = __syncthreads_or(bad);      // NaN / Inf inputs read as "converged": report them (see write_result)
    if (tid == 0) {
        int st = lifted ? DDW_STATUS_REGULARIZED : 0;
        if (nonfinite) { st |= DDW_STATUS_INEXACT; atomicAdd(p.bad_count, 1); }
        else if (!converged) {
            st |= DDW_STATUS_FALLBACK | DDW_STATUS_INEXACT;
            const int sl = atomicAdd(p.fail_count, 1);
            p.fail_list[sl] = (int)b;
        }
        p.status[b] = st;
        if (p.saved_n) p.saved_n[b] = -1;       // this kernel keeps no factor: the backward takes its sparse-support path
    }
}

#ifdef DDW_PHASE_TIMING
}  // namespace ddw
// debug only: copy (and optionally clear) the phase table of this translation unit's kernels
extern "C" int DDW_PHASE_EXPORT(unsigned long long *out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, ddw::ddw_phase_cycles, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(ddw::ddw_phase_cycles, z, sizeof(z)); }
    return 0;
}
namespace ddw {
#endif

template <typename T> int launch_fwd_sparse(const MkFwdParams<T> &p, cudaStream_t st) {
    auto ks = markowitz_fwd_sparse_kernel<T, 128, 4>;
    static DeviceOnce configured;     // per template instance AND per device (the attribute belongs to the device)
    const int dev = current_device();
    if (configured.needs(dev)) {
        if (cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, device_smem_limit()) != cudaSuccess)
            return check_launch("cudaFuncSetAttribute(markowitz_fwd_sparse)");
        configured.mark(dev);
    }
    size_t smem = sp_smem_bytes(p.N, sizeof(T));
#ifdef DDW_PHASE_TIMING
    if (const char *pad = getenv("DDW_DEBUG_SMEM_PAD")) smem += (size_t)atoi(pad);   // fewer resident CTAs per SM
#endif
    ks<<<(unsigned)p.B, 128, smem, st>>>(p, nullptr, nullptr);
    return check_launch("markowitz_fwd_sparse_kernel");
}

// list mode: at most `max_entries` portfolios named by in_list[0 .. *in_count)
template <typename T> int launch_fwd_sparse_list(const MkFwdParams<T> &p, const int *in_list, const int *in_count, int max_entries,
                                                 cudaStream_t st) {
    auto ks = markowitz_fwd_sparse_kernel<T, 128, 4>;
    static DeviceOnce configured;
    const int dev = current_device();
    if (configured.needs(dev)) {
        if (cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, device_smem_limit()) != cudaSuccess)
            return check_launch("cudaFuncSetAttribute(markowitz_fwd_sparse, list mode)");
        configured.mark(dev);
    }
    ks<<<(unsigned)max_entries, 128, sp_smem_bytes(p.N, sizeof(T)), st>>>(p, in_list, in_count);
    return check_launch("markowitz_fwd_sparse_kernel (list mode)");
}

}  // namespace ddw
