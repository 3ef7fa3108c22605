// Standard headers of the device code. Under NVRTC (runtime compilation of expression limit
// states and sampling plans, ebn_jit.cpp) there is no host include tree: the fixed-width integer
// types and the one trait the kernels use are declared here instead.
#pragma once
#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
namespace std {
template <class A, class B>
struct is_same {
  static constexpr bool value = false;
};
template <class A>
struct is_same<A, A> {
  static constexpr bool value = true;
};
}  // namespace std
#else
#include <cstddef>
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>
#endif
