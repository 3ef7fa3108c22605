// Kernel instantiations for the MinAffine limit state.
#include "ebn_mc_impl.cuh"

namespace ebn {

LsVTable vt_min_affine() { return LsLaunch<MinAffine>::vtable(); }
}  // namespace ebn
