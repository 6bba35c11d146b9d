// built-in system "nbody": kernel instantiations (see sde_builtin.cuh)
#include "sde_builtin.cuh"
SDE_DEFINE_BUILTIN(nbody, sde::NBodyLite)
