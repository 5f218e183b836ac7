// markowitz_bwd_f64.cu -- double instantiation of the NumericalMarkowitz backward (markowitz_impl.cuh).
// One translation unit per (direction, dtype) so that nvcc compiles them in parallel.
#include "markowitz_impl.cuh"

namespace ddw {
template int markowitz_bwd_t<double>(const void *, const void *, const int8_t *, const void *, const void *,
                                    const void *, long long, int, void *, void *, void *, void *, void *, size_t,
                                    cudaStream_t, long long, long long, int, const void *, void *);
}  // namespace ddw
