// step_trt.cu — the fused step kernels instantiated for the TRT collision operator
// (d2q9.cuh: collide<COL_TRT>), float and double.  See step_impl.cuh.
#include "step_impl.cuh"

namespace chemsim {
CHEMSIM_INSTANTIATE_STEP(COL_TRT)
CHEMSIM_INSTANTIATE_STEP2(COL_TRT)
}  // namespace chemsim
