// markowitz_bwd_sparse_f32.cu -- float instantiation of the sparse-support NumericalMarkowitz backward kernel
// (markowitz_bwd_sparse.cuh).  Its own translation unit: the hot kernel rebuilds in seconds.
#define DDW_MK_DEFINE_ELIGIBLE
#define DDW_PHASE_EXPORT ddw_debug_bwd_phases_f32
#include "markowitz_bwd_sparse.cuh"

namespace ddw {
template int launch_bwd_sparse<float>(const MkBwdParams<float> &, cudaStream_t);
}  // namespace ddw
