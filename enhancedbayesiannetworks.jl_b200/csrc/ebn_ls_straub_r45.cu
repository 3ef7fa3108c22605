// Kernel instantiations for the Straub frame with observed R4, R5 and the single capacity.
#include "ebn_mc_impl.cuh"

namespace ebn {
LsVTable vt_straub_r45() { return LsLaunch<StraubFrameR45>::vtable(); }
LsVTable vt_straub_capacity() { return LsLaunch<StraubCapacity>::vtable(); }
}  // namespace ebn
