// tc_probe.cu -- standalone check of the tcgen05 tile GEMM (deepdow_b200/csrc/tc_gemm.cuh) against a float64 CPU product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I deepdow_b200/csrc tools/tc_probe.cu -o tools/bin/tc_probe
// Run on a B200: tools/bin/tc_probe [--swap]   (--swap exchanges the LBO / SBO fields of the shared-memory descriptors)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "tc_gemm.cuh"

using namespace ddw;

struct StoreT {   // C[n][m] (transposed, coalesced)
    float *C; int M, N, ldc;
    __device__ void operator()(int m, int n, float v) const { if (m < M && n < N) C[(long long)n * ldc + m] = v; }
};

template <int BN, typename LA, typename LB>
__global__ void __launch_bounds__(tc::kThreads) probe_kernel(LA la, LB lb, int K, StoreT st, uint32_t lbo, uint32_t sbo) {
    extern __shared__ unsigned char smem[];
    __shared__ tc::TileShared sh;
    tc::tile_gemm<BN>(la, lb, blockIdx.x * tc::kTileM, blockIdx.y * BN, K, st, smem, sh, lbo, sbo);
}

static double run_case(const char *name, int M, int N, int K, bool a_rows_fast, bool b_rows_fast, int BN, uint32_t lbo, uint32_t sbo,
                       bool onex) {
    std::vector<float> hA((size_t)M * K), hB((size_t)N * K);
    srand(1234 + M + 7 * N + 13 * K);
    for (auto &v : hA) v = (float)rand() / RAND_MAX - 0.5f;
    for (auto &v : hB) v = (float)rand() / RAND_MAX - 0.5f;
    // memory images: K-contiguous: p[r*K + k]; rows-contiguous: p[k*R + r]
    std::vector<float> mA(hA.size()), mB(hB.size());
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) mA[a_rows_fast ? (size_t)k * M + r : (size_t)r * K + k] = hA[(size_t)r * K + k];
    for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) mB[b_rows_fast ? (size_t)k * N + r : (size_t)r * K + k] = hB[(size_t)r * K + k];
    float *dA, *dB, *dC;
    cudaMalloc(&dA, mA.size() * 4); cudaMalloc(&dB, mB.size() * 4); cudaMalloc(&dC, (size_t)M * N * 4);
    cudaMemcpy(dA, mA.data(), mA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, mB.data(), mB.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xff, (size_t)M * N * 4);
    StoreT st{dC, M, N, M};
    dim3 grid((M + 127) / 128, (N + BN - 1) / BN);
    tc::LoadKContig ak{dA, K, M, K}, bk{dB, K, N, K};
    tc::LoadRowsContig ar{dA, M, M, K, nullptr, nullptr}, br{dB, N, N, K, nullptr, nullptr};
#define LAUNCH(BNV, LA, LB)                                                                                         \
    do {                                                                                                            \
        auto kern = probe_kernel<BNV, decltype(LA), decltype(LB)>;                                                  \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::tile_gemm_smem<BNV>());     \
        kern<<<grid, tc::kThreads, tc::tile_gemm_smem<BNV>()>>>(LA, LB, K, st, lbo, sbo);                            \
    } while (0)
#define PICK(BNV)                                                                   \
    do {                                                                            \
        if (a_rows_fast && b_rows_fast) LAUNCH(BNV, ar, br);                         \
        else if (a_rows_fast) LAUNCH(BNV, ar, bk);                                   \
        else if (b_rows_fast) LAUNCH(BNV, ak, br);                                   \
        else LAUNCH(BNV, ak, bk);                                                    \
    } while (0)
    if (BN == 128) PICK(128); else PICK(64);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(3); }
    std::vector<float> hC((size_t)M * N);
    cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)hA[(size_t)m * K + k] * (double)hB[(size_t)n * K + k];
            const double got = hC[(size_t)n * M + m];
            const double err = std::isfinite(got) ? fabs(got - ref) : 1e30;
            maxerr = err > maxerr ? err : maxerr;
            maxref = fabs(ref) > maxref ? fabs(ref) : maxref;
        }
    printf("%-28s M=%d N=%d K=%d BN=%d  max|err|=%.3e  max|ref|=%.3e  rel=%.3e\n", name, M, N, K, BN, maxerr, maxref, maxerr / maxref);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    (void)onex;
    return maxerr / maxref;
}

int main(int argc, char **argv) {
    uint32_t lbo = tc::kLbo, sbo = tc::kSbo;
    if (argc > 1 && !strcmp(argv[1], "--swap")) { lbo = tc::kSbo; sbo = tc::kLbo; }
    printf("tc_probe: lbo=%u sbo=%u\n", lbo, sbo);
    double worst = 0;
    worst = fmax(worst, run_case("k-contig x k-contig", 128, 128, 32, false, false, 128, lbo, sbo, false));
    worst = fmax(worst, run_case("k-contig x k-contig", 128, 128, 128, false, false, 128, lbo, sbo, false));
    worst = fmax(worst, run_case("rows x rows (Gram form)", 256, 256, 250, true, true, 128, lbo, sbo, false));
    worst = fmax(worst, run_case("tails", 200, 136, 77, false, true, 128, lbo, sbo, false));
    worst = fmax(worst, run_case("BN=64", 130, 64, 128, true, false, 64, lbo, sbo, false));
    printf("worst relative error %.3e -> %s\n", worst, worst < 6e-7 ? "PASS" : "FAIL");
    return worst < 6e-7 ? 0 : 1;
}
