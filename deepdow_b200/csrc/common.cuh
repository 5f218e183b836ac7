// common.cuh -- shared device helpers for the sm_100a kernels of libddw_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "ddw.h"

namespace ddw {

constexpr unsigned kFullMask = 0xffffffffu;

template <typename T> struct Num;
template <> struct Num<float> {
    static __device__ __forceinline__ float eps() { return 1.1920929e-7f; }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float tiny() { return 1e-30f; }
};
template <> struct Num<double> {
    static __device__ __forceinline__ double eps() { return 2.220446049250313e-16; }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ double tiny() { return 1e-300; }
};

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}
template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { T t = __shfl_xor_sync(kFullMask, v, o); v = t > v ? t : v; }
    return v;
}
template <typename T> __device__ __forceinline__ T warp_min(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { T t = __shfl_xor_sync(kFullMask, v, o); v = t < v ? t : v; }
    return v;
}
// (value, index) arg-min over a warp; ties resolved towards the smaller index
template <typename T> __device__ __forceinline__ void warp_argmin(T &v, int &i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T tv = __shfl_xor_sync(kFullMask, v, o);
        int ti = __shfl_xor_sync(kFullMask, i, o);
        if (tv < v || (tv == v && ti < i)) { v = tv; i = ti; }
    }
}

// block-wide sum; `red` is a shared array of at least kThreads/32 elements.  Two barriers.
template <typename T, int kThreads> __device__ __forceinline__ T block_sum(T v, T *red) {
    constexpr int kWarps = kThreads / 32;
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    T s = 0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) s += red[i];
    __syncthreads();
    return s;
}

template <typename T> __device__ __forceinline__ T clampT(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }
template <typename T> __device__ __forceinline__ T absT(T v) { return v < 0 ? -v : v; }

// correctly rounded reciprocal: one MUFU.RCP plus a fix-up for float instead of the generic division sequence
__device__ __forceinline__ float recipT(float v) { return __frcp_rn(v); }
__device__ __forceinline__ double recipT(double v) { return 1.0 / v; }
// reciprocal without the denormal/overflow side path of __frcp_rn: hardware approximation + one Newton step
// (for pivots of an SPD matrix that were already lifted to dmin; error below 1 ulp)
__device__ __forceinline__ float recipFastT(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return fmaf(r, fmaf(-v, r, 1.0f), r);
}
__device__ __forceinline__ double recipFastT(double v) { return 1.0 / v; }

__device__ __forceinline__ float rsqrtT(float v) { return rsqrtf(v); }
__device__ __forceinline__ double rsqrtT(double v) { return rsqrt(v); }
__device__ __forceinline__ float sqrtT(float v) { return sqrtf(v); }
__device__ __forceinline__ double sqrtT(double v) { return sqrt(v); }
__device__ __forceinline__ float expT(float v) { return expf(v); }
__device__ __forceinline__ double expT(double v) { return exp(v); }

// 4 consecutive elements, 16-byte aligned for float / 32-byte for double
template <typename T> struct Vec4 { T v[4]; };
__device__ __forceinline__ Vec4<float> ld4(const float *p) {
    float4 t = *reinterpret_cast<const float4 *>(p);
    Vec4<float> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
__device__ __forceinline__ Vec4<double> ld4(const double *p) {
    double2 a = *reinterpret_cast<const double2 *>(p);
    double2 b = *reinterpret_cast<const double2 *>(p + 2);
    Vec4<double> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; return r;
}

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk): one thread arms the barrier with the byte count and
// issues the copy; every consumer waits on the barrier's phase parity ---------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes,
                                             unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a copy that never lands traps instead of hanging the device
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    for (int spin = 0; spin < (1 << 22); ++spin)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}

__host__ __device__ __forceinline__ int round_up(int v, int m) { return (v + m - 1) / m * m; }

// host-side error plumbing (api.cu)
void set_error(const char *fmt, ...);
int check_launch(const char *what);

// ---- per-device launch configuration ------------------------------------------------------------------------------
// cudaFuncSetAttribute (the > 48 KB dynamic shared-memory opt-in) and the opt-in limit itself belong to a DEVICE: a
// process that drives several GPUs (DataParallel threads, a model moved to cuda:1) must configure every kernel once
// per device it launches on.  DeviceOnce is a per-kernel-instance table of "configured on device d" flags; setting the
// attribute twice is harmless, so plain relaxed atomics are enough for thread safety.
constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev < 0 || dev >= kMaxDevices ? 0 : dev;
}
struct DeviceOnce {
    volatile unsigned char done[kMaxDevices];
    bool needs(int dev) const { return !done[dev]; }
    void mark(int dev) { done[dev] = 1; }
};
// the device's opt-in shared-memory limit per block, cached per device
inline int device_smem_limit() {
    static volatile int lim[kMaxDevices];
    const int dev = current_device();
    int v = lim[dev];
    if (v <= 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) v = 48 * 1024;
        cudaGetLastError();
        lim[dev] = v;
    }
    return v;
}

}  // namespace ddw
