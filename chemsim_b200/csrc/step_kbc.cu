// step_kbc.cu — the fused step kernels instantiated for the KBC collision operator
// (d2q9.cuh: collide<COL_KBC>), float and double.  See step_impl.cuh.
#include "step_impl.cuh"

namespace chemsim {
CHEMSIM_INSTANTIATE_STEP(COL_KBC)
}  // namespace chemsim
