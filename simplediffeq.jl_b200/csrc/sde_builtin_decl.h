// sde_builtin_decl.h -- host-side declarations of the per-system kernel tables
// (defined by SDE_DEFINE_BUILTIN in sde_sys_<name>.cu).
#pragma once

namespace sde {
struct KernelInfo {
  const void* fn;   // device function symbol (usable with cudaLaunchKernel)
  bool adaptive;
};
}  // namespace sde

// variant: bit 0 = reference-exact fixed-step Vern9 dense output (Q2), bit 1 = strict controller,
//          bit 2 = shared-memory staged trajectory-major series output (fixed step)
typedef sde::KernelInfo (*sde_builtin_lookup_fn)(int alg, int dtype, int save, int variant);

sde::KernelInfo sde_lookup_lorenz(int, int, int, int);
sde::KernelInfo sde_lookup_vanderpol(int, int, int, int);
sde::KernelInfo sde_lookup_robertson(int, int, int, int);
sde::KernelInfo sde_lookup_nbody(int, int, int, int);
sde::KernelInfo sde_lookup_lineardecay(int, int, int, int);
sde::KernelInfo sde_lookup_scalargrowth(int, int, int, int);
sde::KernelInfo sde_lookup_nonautonomous(int, int, int, int);
// SDE_COMPAT_FAST_RHS twins (contracted right-hand sides)
sde::KernelInfo sde_lookup_lorenz_fma(int, int, int, int);
sde::KernelInfo sde_lookup_vanderpol_fma(int, int, int, int);
