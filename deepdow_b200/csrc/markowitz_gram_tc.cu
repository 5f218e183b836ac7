// markowitz_gram_tc.cu -- Q = A'A + |alpha| I of NumericalMarkowitz on tensor cores, for universes whose Q lives in the global
// workspace (float32, n_assets beyond one SM's shared memory: BASELINE config 5, n_assets = 500).
//
// A[k][i] = gamma_sqrt[(b N^2 + k N + i) % n_samples] * covmat_sqrt[b][k][i] is the reference's tiled scaling
// (deepdow/layers/allocate.py:268-273); the contraction runs over the ROWS k of A (it is A'A, not AA').  There the Gram is
// 2 N^3 = 250 MFLOP per portfolio and was ~2/3 of the dense forward kernel on FP32 FMA; as a 3xTF32 tcgen05 tile GEMM
// (tc_gemm.cuh) it is a pre-pass of a few milliseconds per 1024 portfolios that leaves Q (lower triangle, row stride LD) where
// markowitz_fwd_kernel expects it.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace ddw {

struct LoadGammaRows {      // element (i, k) = C[k][i] * gamma[(g0 + k N + i) % Bmod]; lanes walk i (coalesced)
    static constexpr bool kRowsFast = true;
    const float *C, *gamma;
    long long g0, Bmod;
    int N;
    __device__ __forceinline__ void ld4(int i, int k, float (&v)[4]) const {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float x = 0.f;
            if (i < N && k + q < N) {
                const long long e = (long long)(k + q) * N + i;
                x = __ldg(C + e) * __ldg(gamma + (g0 + e) % Bmod);
            }
            v[q] = x;
        }
    }
};

struct StoreQLower {
    float *Q; int N, LD; float aabs;
    __device__ __forceinline__ void operator()(int m, int n, float v) const {
        if (m <= n && n < N) Q[(long long)n * LD + m] = m == n ? v + aabs : v;      // Q is symmetric: (n, m) = (m, n)
    }
};

__global__ void __launch_bounds__(tc::kThreads) markowitz_gram_tc_kernel(const float *C, const float *gamma, const float *alpha,
                                                                         long long Bmod, long long b0, int N, int LD, float *Q) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ tc::TileShared sh;
    if (blockIdx.x > blockIdx.y) return;              // only the lower triangle (rows n >= columns m) is consumed
    const long long b = blockIdx.z;
    LoadGammaRows la{C + b * (long long)N * N, gamma, (b0 + b) * (long long)N * N, Bmod, N};
    StoreQLower epi{Q + b * (long long)N * LD, N, LD, fabsf(alpha[b])};
    tc::tile_gemm<128>(la, la, blockIdx.x * 128, blockIdx.y * 128, N, epi, smem_dyn, sh);
}

int markowitz_gram_tc(const float *C, const float *gamma, const float *alpha, long long B, long long Bmod, long long b0, int N,
                      int LD, float *Q, cudaStream_t st) {
    static DeviceOnce configured;
    const int dev = current_device();
    if (configured.needs(dev)) {
        if (cudaFuncSetAttribute(markowitz_gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::tile_gemm_smem<128>()) != cudaSuccess)
            return check_launch("cudaFuncSetAttribute(markowitz_gram_tc_kernel)");
        configured.mark(dev);
    }
    const unsigned t = (unsigned)((N + 127) / 128);
    for (long long z0 = 0; z0 < B; z0 += 32768) {       // grid.z limit
        const long long nb = B - z0 < 32768 ? B - z0 : 32768;
        markowitz_gram_tc_kernel<<<dim3(t, t, (unsigned)nb), tc::kThreads, tc::tile_gemm_smem<128>(), st>>>(
            C + z0 * (long long)N * N, gamma, alpha + z0, Bmod, b0 + z0, N, LD, Q + z0 * (long long)N * LD);
    }
    return check_launch("markowitz_gram_tc_kernel");
}

}  // namespace ddw
