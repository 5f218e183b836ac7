// markowitz_bwd_sparse_f64.cu -- double instantiation of the sparse-support NumericalMarkowitz backward kernel
// (markowitz_bwd_sparse.cuh).  Its own translation unit: the hot kernel rebuilds in seconds.

#define DDW_PHASE_EXPORT ddw_debug_bwd_phases_f64
#include "markowitz_bwd_sparse.cuh"

namespace ddw {
template int launch_bwd_sparse<double>(const MkBwdParams<double> &, cudaStream_t);
}  // namespace ddw
