// Kernel instantiations for the PolynomialProgram limit state.
#include "ebn_mc_impl.cuh"

namespace ebn {

LsVTable vt_polynomial() { return LsLaunch<PolynomialProgram>::vtable(); }
}  // namespace ebn
