// markowitz_fwd_sparse_f64.cu -- double instantiation of the sparse-support NumericalMarkowitz forward kernel
// (markowitz_sparse.cuh).  Its own translation unit: the hot kernel rebuilds in seconds.
#define DDW_PHASE_EXPORT ddw_debug_fwd_phases_f64
#include "markowitz_sparse.cuh"

namespace ddw {
template int launch_fwd_sparse<double>(const MkFwdParams<double> &, cudaStream_t);
}  // namespace ddw
