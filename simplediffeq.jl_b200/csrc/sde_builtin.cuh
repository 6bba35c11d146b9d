// sde_builtin.cuh -- __global__ wrappers + (alg, dtype, save) -> kernel table for one built-in system.
// Each sde_sys_<name>.cu instantiates the table for its system so the systems compile in parallel.
#pragma once
#include "device/sde_kernels.cuh"
#include "device/sde_systems.cuh"
#include "sde_builtin_decl.h"

#ifndef SDE_BLOCK
#define SDE_BLOCK 128
#endif

namespace sde {

// staged kernels: four CTAs per SM is what their shared-memory footprint allows (StageCfg) -- hold ptxas to the
// 128 registers that go with it, except where the stage vectors alone need more (FP64 Verner methods, large states)
template <class Method, class T, int N, bool STAGED>
struct FixedMinBlocks {
  static constexpr bool kHeavy = (Method::kNB > 7 && sizeof(T) == 8) || N > 4;
  static constexpr int value = (STAGED && !kHeavy) ? 4 : 1;
};
template <class Sys, class T, class Method, int SAVE, bool Q2, bool STAGED>
__global__ void __launch_bounds__(SDE_BLOCK, FixedMinBlocks<Method, T, Sys::N, STAGED>::value) fixed_kernel(const __grid_constant__ KArgs<T> a) {
  fixed_body<Sys, T, Method, SAVE, Q2, STAGED>(a);
}

template <class Sys, class T, class Method, int SAVE, bool kV9, bool kStrict>
__global__ void __launch_bounds__(SDE_BLOCK) adaptive_kernel(const __grid_constant__ KArgs<T> a) {
  adaptive_body<Sys, T, Method, SAVE, kV9, kStrict>(a);
}

// staged variants exist only for the series modes
template <class Sys, class T, class M, int S, bool Q>
inline KernelInfo pick_fixed(bool staged) {
  if constexpr (S != kSaveEndpoint) {
    if (staged) return KernelInfo{(const void*)&fixed_kernel<Sys, T, M, S, Q, true>, false};
  }
  return KernelInfo{(const void*)&fixed_kernel<Sys, T, M, S, Q, false>, false};
}

template <class Sys, class T>
inline KernelInfo lookup_kernel_t(int alg, int save, bool q2, bool strict, bool staged, bool fast_stages) {
  using TS = Tsit5Method<Sys, T>;
  using TF = Tsit5FastMethod<Sys, T>;     // SDE_COMPAT_FAST_STAGES: the step size folded into the stage coefficients
  using RK = RK4Method<Sys, T>;
  using EU = EulerMethod<Sys, T>;
  using V7 = Vern7Method<Sys, T>;
  using V9 = Vern9Method<Sys, T>;
#define SDE_FIXED(M, S, Q) pick_fixed<Sys, T, M, S, Q>(staged)
#define SDE_ADAPT(M, S, V)                                                       \
  (strict ? KernelInfo{(const void*)&adaptive_kernel<Sys, T, M, S, V, true>, true} \
          : KernelInfo{(const void*)&adaptive_kernel<Sys, T, M, S, V, false>, true})
  switch (alg) {
    case kTsit5:
      if (save == kSaveEndpoint) return fast_stages ? SDE_FIXED(TF, kSaveEndpoint, false) : SDE_FIXED(TS, kSaveEndpoint, false);
      if (save == kSaveAt) return fast_stages ? SDE_FIXED(TF, kSaveAt, false) : SDE_FIXED(TS, kSaveAt, false);
      if (save == kSaveEveryStep) return fast_stages ? SDE_FIXED(TF, kSaveEveryStep, false) : SDE_FIXED(TS, kSaveEveryStep, false);
      break;
    case kRK4:   // the reference's GPUSimpleRK4 has no saveat
      if (save == kSaveEndpoint) return SDE_FIXED(RK, kSaveEndpoint, false);
      if (save == kSaveEveryStep) return SDE_FIXED(RK, kSaveEveryStep, false);
      break;
    case kEuler:  // like RK4: no saveat in the reference
      if (save == kSaveEndpoint) return SDE_FIXED(EU, kSaveEndpoint, false);
      if (save == kSaveEveryStep) return SDE_FIXED(EU, kSaveEveryStep, false);
      break;
    case kVern7:
      if (save == kSaveEndpoint) return SDE_FIXED(V7, kSaveEndpoint, false);
      if (save == kSaveAt) return SDE_FIXED(V7, kSaveAt, false);
      if (save == kSaveEveryStep) return SDE_FIXED(V7, kSaveEveryStep, false);
      break;
    case kVern9:
      if (save == kSaveEndpoint) return SDE_FIXED(V9, kSaveEndpoint, false);
      if (save == kSaveAt) return q2 ? SDE_FIXED(V9, kSaveAt, true) : SDE_FIXED(V9, kSaveAt, false);
      if (save == kSaveEveryStep) return SDE_FIXED(V9, kSaveEveryStep, false);
      break;
    case kATsit5:
      if (save == kSaveEndpoint) return SDE_ADAPT(TS, kSaveEndpoint, false);
      if (save == kSaveAt) return SDE_ADAPT(TS, kSaveAt, false);
      if (save == kSaveEveryStep) return SDE_ADAPT(TS, kSaveEveryStep, false);
      break;
    case kAVern7:
      if (save == kSaveEndpoint) return SDE_ADAPT(V7, kSaveEndpoint, false);
      if (save == kSaveAt) return SDE_ADAPT(V7, kSaveAt, false);
      if (save == kSaveEveryStep) return SDE_ADAPT(V7, kSaveEveryStep, false);
      break;
    case kAVern9:
      if (save == kSaveEndpoint) return SDE_ADAPT(V9, kSaveEndpoint, true);
      if (save == kSaveAt) return SDE_ADAPT(V9, kSaveAt, true);
      if (save == kSaveEveryStep) return SDE_ADAPT(V9, kSaveEveryStep, true);
      break;
  }
#undef SDE_FIXED
#undef SDE_ADAPT
  return KernelInfo{nullptr, false};
}

template <class Sys>
inline KernelInfo lookup_kernel(int alg, int dtype, int save, int variant) {
  const bool q2 = (variant & 1) != 0, strict = (variant & 2) != 0, staged = (variant & 4) != 0, fast_stages = (variant & 8) != 0;
  return dtype == 0 ? lookup_kernel_t<Sys, double>(alg, save, q2, strict, staged, fast_stages)
                    : lookup_kernel_t<Sys, float>(alg, save, q2, strict, staged, fast_stages);
}

}  // namespace sde

#define SDE_DEFINE_BUILTIN(name, SysT)                                                        \
  sde::KernelInfo sde_lookup_##name(int alg, int dtype, int save, int variant) {              \
    return sde::lookup_kernel<SysT>(alg, dtype, save, variant);                               \
  }
