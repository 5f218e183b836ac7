// markowitz_fwd_sparse_f32.cu -- float instantiation of the sparse-support NumericalMarkowitz forward kernel
// (markowitz_sparse.cuh).  Its own translation unit: the hot kernel rebuilds in seconds.
#define DDW_MK_DEFINE_ELIGIBLE
#define DDW_PHASE_EXPORT ddw_debug_fwd_phases_f32
#include "markowitz_sparse.cuh"

namespace ddw {
template int launch_fwd_sparse<float>(const MkFwdParams<float> &, cudaStream_t);
template int launch_fwd_sparse_list<float>(const MkFwdParams<float> &, const int *, const int *, int, cudaStream_t);
}  // namespace ddw
