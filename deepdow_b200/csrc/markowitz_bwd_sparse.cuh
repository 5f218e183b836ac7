// markowitz_bwd_sparse.cuh -- backward of NumericalMarkowitz for portfolios with a small free set
// (compiled in markowitz_bwd_sparse_f32.cu / markowitz_bwd_sparse_f64.cu; formulas in markowitz_impl.cuh:
// implicit differentiation of the KKT system on the active set saved by the forward)
//
//   F = {state == 0};  [[2 Q_FF, 1],[1', 0]] [d_F; eta] = [g_F; 0],  d = 0 off F
//   d_rets = d;  dA = -2 [(A d) w' + (A w) d'];  d_covmat_sqrt = gamma_ * dA;  d_alpha = sign(alpha)(-2 d.w)
//
// One CTA of 128 threads per portfolio.  A (gamma-tiled covmat_sqrt) is brought into shared memory by one bulk
// TMA copy; Q_FF is formed from the free columns as 4x4 register tiles, the reduced KKT system is factored by all
// four warps with every element in a register, A d and A w are row passes over the support, and the rank-2 gradient is streamed out with 16-byte stores
// (gamma loads issued one step ahead).  Portfolios with more than kCandMax free assets go to the dense kernel.
#pragma once
#include "markowitz_common.cuh"
#include "support_kernels.cuh"
#include "warp_solve.cuh"

namespace ddw {

constexpr int kBwdFreeMax = 48;
constexpr int kBwdLd = (kBwdFreeMax + 3) | 1;
constexpr int kBwdSlots = (kBwdFreeMax + 31) / 32;

constexpr int bs_round_up(int v, int m) { return (v + m - 1) / m * m; }
constexpr size_t bs_smem_bytes(int N, int tsize) {
    return 16 + ((size_t)bs_round_up(N * N, 4) + bs_round_up(kBwdFreeMax * kBwdLd, 4) + 8 * (size_t)bs_round_up(N + 2, 32) + 40) * tsize +
           ((size_t)bs_round_up(N + 2, 32) + 8) * sizeof(int);
}
static_assert(4 * (bs_smem_bytes(100, 4) + 1024) <= 228 * 1024, "sparse backward kernel: n_assets=100 must fit 4 CTAs per SM");

#ifdef DDW_MK_DEFINE_ELIGIBLE
bool bwd_sparse_eligible(int N, int tsize, int smem_limit) {
    return N <= 128 && N >= 8 && bs_smem_bytes(N, tsize) <= (size_t)smem_limit / 2;
}
#else
bool bwd_sparse_eligible(int N, int tsize, int smem_limit);
#endif

template <typename T, int kThreads>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? 4 : 2) markowitz_bwd_sparse_kernel(MkBwdParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long long b = blockIdx.x;
    const int N = p.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int LDC = kBwdLd;
    const int NP = round_up(N + 2, 32);
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw);
    T *ptr = reinterpret_cast<T *>(smem_raw + 16);
    T *As = ptr; ptr += round_up(N * N, 4);
    T *Lb = ptr; ptr += round_up(kBwdFreeMax * kBwdLd, 4);
    T *wv = ptr; ptr += NP;
    T *dv = ptr; ptr += NP;
    T *gv = ptr; ptr += NP;
    T *pv = ptr; ptr += NP;
    T *qv = ptr; ptr += NP;
    T *dinv = ptr; ptr += NP;
    T *wS = ptr; ptr += NP;      // w and d on the compacted support (free assets first, then capped)
    T *dS = ptr; ptr += NP;
    T *red = ptr; ptr += 40;
    int *idx = reinterpret_cast<int *>(ptr);
    int *misc = idx + NP;

    const T *Cb = p.C + b * (long long)N * N;
    const long long gflat0 = (p.b0 + b) * (long long)N * N;
    DDW_PHASE_INIT
    DDW_PHASE_COUNT(11);
    const bool use_tma = stage_matrix_begin<T, kThreads>(As, Cb, N, mbar);
    // support list from the state saved by the forward: free assets first (ascending), then the capped ones
    if (warp == 0) {
        int nf = 0, nu = 0;
        for (int base = 0; base < N; base += 32) {
            const int i = base + lane;
            const int v = i < N ? (int)p.state[b * N + i] : 2;
            const unsigned mf = __ballot_sync(kFullMask, v == 0);
            if (v == 0) idx[nf + __popc(mf & ((1u << lane) - 1u))] = i;
            nf += __popc(mf);
        }
        for (int base = 0; base < N; base += 32) {
            const int i = base + lane;
            const int v = i < N ? (int)p.state[b * N + i] : 2;
            const unsigned mu = __ballot_sync(kFullMask, v == 1);
            if (v == 1 && nf + nu + __popc(mu & ((1u << lane) - 1u)) < NP) idx[nf + nu + __popc(mu & ((1u << lane) - 1u))] = i;
            nu += __popc(mu);
        }
        if (lane == 0) { misc[0] = nf; misc[1] = nu; }
    }
    for (int i = tid; i < N; i += kThreads) {
        wv[i] = p.w[b * N + i];
        gv[i] = p.grad_w[b * N + i];
        dv[i] = T(0);
    }
    const T alpha = p.alpha[b];
    stage_matrix_end<T, kThreads>(As, Cb, N, mbar, use_tma);
    DDW_PHASE(0);
    apply_gamma_tiling<T, kThreads>(As, N, p.gamma, gflat0, p.Bmod);
    __syncthreads();
    DDW_PHASE(1);
    const int nF = misc[0], nS = misc[0] + misc[1];
    if (nF > kBwdFreeMax) {
        // free set too large for this kernel: hand the portfolio to the dense backward
        if (tid == 0) { const int sl = atomicAdd(p.work_count, 1); p.work_list[sl] = (int)b; }
        return;
    }
    for (int m = tid; m < NP; m += kThreads) { wS[m] = m < nS ? wv[idx[m]] : T(0); dS[m] = T(0); }
    if (nF > 0) {
        // Q_FF = A_F' A_F + |alpha| I straight into the factor area (element (s,c) at Lb[c*LDC + s])
        gram_columns<T, kThreads>(As, N, idx, (const T *)nullptr, absT(alpha), 0, nF, Lb, 1, LDC);
        __syncthreads();
        DDW_PHASE(2);
        bool solved = false;
        if constexpr (sizeof(T) == 4) {
            if (nF <= kWsFrame) {
                // at most 32 free assets (float): warp 0 solves the reduced KKT system in registers (warp_solve.cuh) -- no block
                // barrier per column, ~2.5 k cycles instead of ~10 k for the four-warp factorisation + one-warp finish
                solved = true;
                if (warp == 0) {
                    const int k0 = kWsFrame - nF, ii = lane - k0;
                    const bool live = lane >= k0;
                    float v[kWsFrame];
#pragma unroll
                    for (int j0 = 0; j0 < kWsFrame; j0 += 4) {
                        if (j0 + 3 >= k0) {
#pragma unroll
                            for (int j = j0; j < j0 + 4; ++j) {
                                const int jj = j - k0;
                                v[j] = (live && jj >= 0) ? (jj <= ii ? Lb[jj * LDC + ii] : Lb[ii * LDC + jj]) : 0.f;
                            }
                        } else {
#pragma unroll
                            for (int j = j0; j < j0 + 4; ++j) v[j] = 0.f;
                        }
                    }
                    float y1 = live ? gv[idx[ii]] : 0.f, y2 = live ? 1.f : 0.f, dinv_me = 0.f;
                    const float dmax_w = warp_max(live ? Lb[ii * LDC + ii] : 0.f);
                    const float dmin_w = 4.f * Num<float>::eps() * dmax_w + Num<float>::tiny();
                    __syncwarp();                      // Q_FF is in registers: its shared array becomes the multiplier array
                    ws_eliminate(v, y1, y2, dinv_me, dmin_w, Lb, lane, k0);
                    const float s12 = warp_sum(y1 * y2 * dinv_me), s22 = warp_sum(y2 * y2 * dinv_me);
                    const float eta = s12 / s22;
                    const float x = ws_backsub(0.5f * (y1 - eta * y2) * dinv_me, Lb, lane, k0);
                    if (live) { dv[idx[ii]] = x; dS[ii] = x; }
                }
            }
        }
        if (!solved) {
        T dmax = T(0);
        for (int i = tid; i < nF; i += kThreads) dmax = max(dmax, Lb[i * LDC + i]);
        dmax = warp_max(dmax);
        if (lane == 0) red[warp] = dmax;
        __syncthreads();
        dmax = T(0);
        for (int i = 0; i < kThreads / 32; ++i) dmax = max(dmax, red[i]);
        const T dmin = T(4) * Num<T>::eps() * dmax + Num<T>::tiny();
        // all four warps factor the augmented triangle [Q_FF; g_F; 1'] (markowitz_common.cuh), warp 0 finishes
        auto init = [&](int i, int j) -> T { return i < nF ? Lb[j * LDC + i] : (i == nF ? gv[idx[j]] : T(1)); };
        if (small_system_fits<T, kThreads>(nF)) factor_small<T, kThreads>(Lb, LDC, nF, dinv, dmin, init);
        else {
            for (int a = tid; a < nF; a += kThreads) { Lb[a * LDC + nF] = gv[idx[a]]; Lb[a * LDC + nF + 1] = T(1); }
            __syncthreads();
            factor_ldl<T, kThreads>(Lb, LDC, nF, 2, dinv, dmin);
        }
        DDW_PHASE(3);
        if (warp == 0) {
            T s12 = T(0), s22 = T(0);
            for (int k = lane; k < nF; k += 32) {
                const T y1 = Lb[k * LDC + nF], y2 = Lb[k * LDC + nF + 1], di = dinv[k];
                s12 = fma(y1 * y2, di, s12);
                s22 = fma(y2 * y2, di, s22);
            }
            s12 = warp_sum(s12); s22 = warp_sum(s22);
            const T eta = s12 / s22;
            T zt[kBwdSlots];
#pragma unroll
            for (int m = 0; m < kBwdSlots; ++m) {
                const int k = lane + 32 * m;
                zt[m] = k < nF ? T(0.5) * (Lb[k * LDC + nF] - eta * Lb[k * LDC + nF + 1]) : T(0);
            }
            backsub_warp<T, kBwdSlots>(Lb, LDC, nF, dinv, zt);
#pragma unroll
            for (int m = 0; m < kBwdSlots; ++m) {
                const int k = lane + 32 * m;
                if (k < nF) { dv[idx[k]] = zt[m]; dS[k] = zt[m]; }
            }
        }
        }   // !solved
    }
    __syncthreads();
    DDW_PHASE(4);
    // d_rets, d_alpha
    T ldot = T(0);
    for (int i = tid; i < N; i += kThreads) { p.d_rets[b * N + i] = dv[i]; ldot = fma(dv[i], wv[i], ldot); }
    const T dot = block_sum<T, kThreads>(ldot, red);
    if (tid == 0) p.d_alpha[b] = (alpha > T(0) ? T(1) : (alpha < T(0) ? T(-1) : T(0))) * (T(-2) * dot);
    // p = A d, q = A w: one thread per row, columns of the support only
    for (int k = tid; k < N; k += kThreads) {
        const T *row = As + k * N;
        T sp = T(0), sq = T(0);
        int m = 0;
        if (nF > 0) {
            for (; m + 3 < nS; m += 4) {
                const int4 j4 = *reinterpret_cast<const int4 *>(idx + m);
                const Vec4<T> w4 = ld4(wS + m), d4 = ld4(dS + m);
                const T a0 = row[j4.x], a1 = row[j4.y], a2 = row[j4.z], a3 = row[j4.w];
                sp = fma(a0, d4.v[0], sp); sq = fma(a0, w4.v[0], sq);
                sp = fma(a1, d4.v[1], sp); sq = fma(a1, w4.v[1], sq);
                sp = fma(a2, d4.v[2], sp); sq = fma(a2, w4.v[2], sq);
                sp = fma(a3, d4.v[3], sp); sq = fma(a3, w4.v[3], sq);
            }
            for (; m < nS; ++m) { const T a = row[idx[m]]; sp = fma(a, dS[m], sp); sq = fma(a, wS[m], sq); }
        }
        pv[k] = sp; qv[k] = sq;      // no free asset: d = 0 and every gradient vanishes
    }
    __syncthreads();
    DDW_PHASE(5);
    if (p.vecs) for (int i = tid; i < N; i += kThreads) { p.vecs[(b * 2) * N + i] = pv[i]; p.vecs[(b * 2 + 1) * N + i] = qv[i]; }
    // d_covmat_sqrt = gamma_ * (-2) (p w' + q d'), flat over the N^2 elements so that every lane stores
    T *dC = p.d_C + b * (long long)N * N;
    const unsigned uB = (unsigned)p.Bmod;
    if ((N & 3) == 0 && (uB & 3u) == 0u) {
        const int n4 = N * N / 4;
        int j = (4 * tid) / N, k = 4 * tid - j * N;
        const int dj = (4 * kThreads) / N, dk = (4 * kThreads) % N;
        unsigned gi = (unsigned)((gflat0 + 4 * tid) % p.Bmod);
        const unsigned gstep = (unsigned)((4 * kThreads) % p.Bmod);
        Vec4<T> g4 = tid < n4 ? ld4(p.gamma + gi) : Vec4<T>();
        for (int e4 = tid; e4 < n4; e4 += kThreads) {
            gi += gstep; if (gi >= uB) gi -= uB;
            Vec4<T> gn = g4;
            if (e4 + kThreads < n4) gn = ld4(p.gamma + gi);    // next step's factors: in flight during this step
            const T pj = T(-2) * pv[j], qj = T(-2) * qv[j];
            const Vec4<T> w4 = ld4(wv + k), d4 = ld4(dv + k);
            Vec4<T> o;
#pragma unroll
            for (int x = 0; x < 4; ++x) o.v[x] = g4.v[x] * fma(pj, w4.v[x], qj * d4.v[x]);
            st4(dC + 4 * e4, o);
            g4 = gn;
            k += dk; j += dj;
            if (k >= N) { k -= N; ++j; }
        }
    } else {
        int j = tid / N, k = tid - j * N;
        const int dj = kThreads / N, dk = kThreads % N;
        unsigned gi = (unsigned)((gflat0 + tid) % p.Bmod);
        const unsigned gstep = (unsigned)(kThreads % p.Bmod);
        for (int e = tid; e < N * N; e += kThreads) {
            dC[e] = p.gamma[gi] * T(-2) * fma(pv[j], wv[k], qv[j] * dv[k]);
            k += dk; j += dj;
            if (k >= N) { k -= N; ++j; }
            gi += gstep; if (gi >= uB) gi -= uB;
        }
    }
    DDW_PHASE(6);
}

#ifdef DDW_PHASE_TIMING
}  // namespace ddw
extern "C" int DDW_PHASE_EXPORT(unsigned long long *out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, ddw::ddw_phase_cycles, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(ddw::ddw_phase_cycles, z, sizeof(z)); }
    return 0;
}
namespace ddw {
#endif

template <typename T> int launch_bwd_sparse(const MkBwdParams<T> &p, cudaStream_t st) {
    auto kb = markowitz_bwd_sparse_kernel<T, 128>;
    static DeviceOnce configured;     // per template instance and per device
    const int dev = current_device();
    if (configured.needs(dev)) {
        if (cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, device_smem_limit()) != cudaSuccess)
            return check_launch("cudaFuncSetAttribute(markowitz_bwd_sparse)");
        configured.mark(dev);
    }
    kb<<<(unsigned)p.B, 128, bs_smem_bytes(p.N, sizeof(T)), st>>>(p);
    return check_launch("markowitz_bwd_sparse_kernel");
}

}  // namespace ddw
