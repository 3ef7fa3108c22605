// Kernel instantiations for the StraubFrame limit state.
#include "ebn_mc_impl.cuh"

namespace ebn {
// Straub: Ur normal, V Gamma, H Gumbel, five in-model standard normals.
LsVTable vt_straub() {
  return LsLaunch<StraubFrame,
                  StaticPlan<DK_NORMAL_BM, DK_GAMMA_MT, DK_GUMBEL, DK_NORMAL_BM, DK_NORMAL_BM, DK_NORMAL_BM, DK_NORMAL_BM, DK_NORMAL_BM>,
                  StaticPlan<DK_NORMAL_BM32, DK_GAMMA_MT, DK_GUMBEL, DK_NORMAL_BM32, DK_NORMAL_BM32, DK_NORMAL_BM32, DK_NORMAL_BM32,
                             DK_NORMAL_BM32>>::vtable();
}
}  // namespace ebn
