// Kernel instantiations for the VehicleSuspension limit state.
#include "ebn_mc_impl.cuh"

namespace ebn {
// vehicle suspension W1/W5: (A, b0) Parameters, V uniform (a truncated uniform is
// again uniform), C, Ck, K untruncated normals.
LsVTable vt_vehicle() { return LsLaunch<VehicleSuspension, StaticPlan<DK_CONST, DK_CONST, DK_UNIFORM, DK_NORMAL_BM, DK_NORMAL_BM, DK_NORMAL_BM>>::vtable(); }
}  // namespace ebn
