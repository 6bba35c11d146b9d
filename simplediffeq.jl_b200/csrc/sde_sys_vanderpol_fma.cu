// SDE_COMPAT_FAST_RHS twin of built-in system "vanderpol": kernel instantiations (see sde_builtin.cuh)
#include "sde_builtin.cuh"
SDE_DEFINE_BUILTIN(vanderpol_fma, sde::VanDerPolFma)
