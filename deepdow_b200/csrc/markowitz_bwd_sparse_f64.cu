// markowitz_bwd_sparse_f64.cu -- double instantiation of the sparse-support NumericalMarkowitz backward kernel
// (markowitz_bwd_sparse.cuh).  Its own translation unit: the hot kernel rebuilds in seconds.

#include "markowitz_bwd_sparse.cuh"

namespace ddw {
template int launch_bwd_sparse<double>(const MkBwdParams<double> &, cudaStream_t);
}  // namespace ddw
