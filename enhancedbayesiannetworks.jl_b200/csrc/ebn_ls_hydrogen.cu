// Kernel instantiations for the HydrogenRadius limit state.
#include "ebn_mc_impl.cuh"

namespace ebn {

LsVTable vt_hydrogen() { return LsLaunch<HydrogenRadius>::vtable(); }
}  // namespace ebn
