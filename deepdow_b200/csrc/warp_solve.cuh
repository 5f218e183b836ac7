// warp_solve.cuh -- a symmetric positive definite system of at most 32 unknowns solved by ONE warp in registers.
//
// Row i of the system lives in lane i; the system sits at the END of a 32 x 32 frame (rows = lanes k0..31, columns = registers
// k0..31, k0 = 32 - n), so that a system of any size runs the tail of one fully unrolled elimination (compile-time register
// indices, entered through a switch) and executes exactly its own triangular share of shuffle + FMA pairs.  Two right-hand sides
// ride along.  The multipliers go to a 32 x 33 shared array for the column-oriented back substitution.  Same scheme as the
// forward's markowitz_qwarp_kernel (markowitz_qwarp.cu), here for the backward of the sparse-support kernel.
#pragma once
#include "common.cuh"

namespace ddw {

constexpr int kWsFrame = 32;
constexpr int kWsLd = 33;

template <int K>
__device__ __forceinline__ void ws_step(float (&v)[kWsFrame], float &y1, float &y2, float &dinv_me, float dmin, float *Lt, int lane) {
    float d = __shfl_sync(kFullMask, v[K], K);
    if (!(d > dmin)) d = dmin;
    const float inv = recipFastT(d);
    const float l = lane > K ? v[K] * inv : 0.f;
    if (lane == K) dinv_me = inv;
    Lt[K * kWsLd + lane] = l;
#pragma unroll
    for (int j = K + 1; j < kWsFrame; ++j) {
        const float rk = __shfl_sync(kFullMask, v[j], K);
        v[j] = fmaf(-l, rk, v[j]);
    }
    const float y1k = __shfl_sync(kFullMask, y1, K), y2k = __shfl_sync(kFullMask, y2, K);
    y1 = fmaf(-l, y1k, y1); y2 = fmaf(-l, y2k, y2);
}

// eliminates rows k0..31 (LDL' without pivoting, pivots lifted to dmin); on return y1, y2 hold L^-1 rhs and dinv_me = 1 / d_lane
__device__ __forceinline__ void ws_eliminate(float (&v)[kWsFrame], float &y1, float &y2, float &dinv_me, float dmin, float *Lt, int lane,
                                             int k0) {
#define DDW_WS_CASE(K) case K: ws_step<K>(v, y1, y2, dinv_me, dmin, Lt, lane);
    switch (k0) {
        DDW_WS_CASE(0) DDW_WS_CASE(1) DDW_WS_CASE(2) DDW_WS_CASE(3) DDW_WS_CASE(4) DDW_WS_CASE(5) DDW_WS_CASE(6) DDW_WS_CASE(7)
        DDW_WS_CASE(8) DDW_WS_CASE(9) DDW_WS_CASE(10) DDW_WS_CASE(11) DDW_WS_CASE(12) DDW_WS_CASE(13) DDW_WS_CASE(14) DDW_WS_CASE(15)
        DDW_WS_CASE(16) DDW_WS_CASE(17) DDW_WS_CASE(18) DDW_WS_CASE(19) DDW_WS_CASE(20) DDW_WS_CASE(21) DDW_WS_CASE(22) DDW_WS_CASE(23)
        DDW_WS_CASE(24) DDW_WS_CASE(25) DDW_WS_CASE(26) DDW_WS_CASE(27) DDW_WS_CASE(28) DDW_WS_CASE(29) DDW_WS_CASE(30) DDW_WS_CASE(31)
        default: break;
    }
#undef DDW_WS_CASE
}

// L' x = z, column by column: lane i holds z_i on entry and x_i on return (rows k0..31)
__device__ __forceinline__ float ws_backsub(float x, const float *Lt, int lane, int k0) {
    __syncwarp();
#pragma unroll 1
    for (int i = kWsFrame - 1; i > k0; --i) {
        const float xi = __shfl_sync(kFullMask, x, i);
        if (lane >= k0 && lane < i) x = fmaf(-Lt[lane * kWsLd + i], xi, x);
    }
    return x;
}

}  // namespace ddw
