// Kernel instantiations for the VehicleSuspension limit state.
#include "ebn_mc_impl.cuh"

namespace ebn {
// vehicle suspension W1/W5: (A, b0) Parameters, V uniform (a truncated uniform is
// again uniform), C, Ck, K untruncated normals.
// Second plan: V a Parameter as well — what the imprecise outer loop (DoubleLoop over an interval V,
// demo/vehicle_suspension_imprecise.jl) and `probability_of_failure` with a fixed V produce.
// Plans 3 and 4: the same two with the normals drawn in fp32 (EBN_JOB_FP32_SAMPLING).
// Plans 5-11: every combination of C, Ck, K truncated normals — what discretised root nodes hand to their
// children (`truncated(dist, a, b)`, reference src/networks/ebn/discretization/discretize.jl:54,
// src/nodes/continuousnodes.jl:71-73): inverse CDF on the truncation window (normcdfinv_fast).
LsVTable vt_vehicle() {
  return LsLaunch<VehicleSuspension,
                  StaticPlan<DK_CONST, DK_CONST, DK_UNIFORM, DK_NORMAL_BM, DK_NORMAL_BM, DK_NORMAL_BM>,
                  StaticPlan<DK_CONST, DK_CONST, DK_CONST, DK_NORMAL_BM, DK_NORMAL_BM, DK_NORMAL_BM>,
                  StaticPlan<DK_CONST, DK_CONST, DK_UNIFORM, DK_NORMAL_BM32, DK_NORMAL_BM32, DK_NORMAL_BM32>,
                  StaticPlan<DK_CONST, DK_CONST, DK_CONST, DK_NORMAL_BM32, DK_NORMAL_BM32, DK_NORMAL_BM32>,
                  StaticPlan<DK_CONST, DK_CONST, DK_UNIFORM, DK_NORMAL_ICDF, DK_NORMAL_BM, DK_NORMAL_BM>,
                  StaticPlan<DK_CONST, DK_CONST, DK_UNIFORM, DK_NORMAL_BM, DK_NORMAL_ICDF, DK_NORMAL_BM>,
                  StaticPlan<DK_CONST, DK_CONST, DK_UNIFORM, DK_NORMAL_BM, DK_NORMAL_BM, DK_NORMAL_ICDF>,
                  StaticPlan<DK_CONST, DK_CONST, DK_UNIFORM, DK_NORMAL_ICDF, DK_NORMAL_ICDF, DK_NORMAL_BM>,
                  StaticPlan<DK_CONST, DK_CONST, DK_UNIFORM, DK_NORMAL_ICDF, DK_NORMAL_BM, DK_NORMAL_ICDF>,
                  StaticPlan<DK_CONST, DK_CONST, DK_UNIFORM, DK_NORMAL_BM, DK_NORMAL_ICDF, DK_NORMAL_ICDF>,
                  StaticPlan<DK_CONST, DK_CONST, DK_UNIFORM, DK_NORMAL_ICDF, DK_NORMAL_ICDF, DK_NORMAL_ICDF>>::vtable();
}
}  // namespace ebn
