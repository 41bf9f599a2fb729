// step_bgk.cu — the fused step kernels instantiated for the BGK collision operator
// (d2q9.cuh: collide<COL_BGK>), float and double.  See step_impl.cuh.
#include "step_impl.cuh"

namespace chemsim {
CHEMSIM_INSTANTIATE_STEP(COL_BGK)
CHEMSIM_INSTANTIATE_STEP2(COL_BGK)
}  // namespace chemsim
