// step_regularized.cu — the fused step kernels instantiated for the REGULARIZED collision operator
// (d2q9.cuh: collide<COL_REGULARIZED>), float and double.  See step_impl.cuh.
#include "step_impl.cuh"

namespace chemsim {
CHEMSIM_INSTANTIATE_STEP(COL_REGULARIZED)
CHEMSIM_INSTANTIATE_STEP2(COL_REGULARIZED)
}  // namespace chemsim
