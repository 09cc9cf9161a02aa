// Logits-in instantiation of the fused loss kernels: the kernel applies ATen's sigmoid itself and
// returns dL/dlogits.  Separate translation unit so the two halves of the template space compile in parallel.
#include "cldet_loss_head_kernels.cuh"

namespace cldet {

template void run_loss_kernels<true>(const LossArgs&, int, bool, bool, bool, dim3, cudaStream_t);
template void run_reweight_kernels<true>(const LossArgs&, int, bool, bool, dim3, cudaStream_t);
template void run_head_loss_kernels<true>(const LossArgs&, const HeadLevels&, bool, bool, bool, dim3, cudaStream_t);
template void run_head_reweight_kernels<true>(const LossArgs&, const HeadLevels&, bool, bool, dim3, cudaStream_t);
template void run_head_flag_kernels<true>(const HeadLevels&, int, int64_t, int, int, uint32_t*, cudaStream_t);

}  // namespace cldet
