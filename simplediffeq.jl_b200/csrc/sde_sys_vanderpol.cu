// built-in system "vanderpol": kernel instantiations (see sde_builtin.cuh)
#include "sde_builtin.cuh"
SDE_DEFINE_BUILTIN(vanderpol, sde::VanDerPol)
