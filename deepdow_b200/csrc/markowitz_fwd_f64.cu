// markowitz_fwd_f64.cu -- double instantiation of the NumericalMarkowitz forward solver (markowitz_impl.cuh).
// One translation unit per (direction, dtype) so that nvcc compiles them in parallel.
#include "markowitz_impl.cuh"

namespace ddw {
template int markowitz_fwd_t<double>(const void *, const void *, const void *, const void *, double, long long, int,
                                    void *, int8_t *, int32_t *, void *, size_t, cudaStream_t, long long, long long, void *);
template int markowitz_fused_fwd_t<double>(const void *, const void *, int, int, const void *, const void *, const void *, double,
                                          long long, int, void *, int8_t *, int32_t *, void *, size_t, cudaStream_t, void *, void *);
}  // namespace ddw
