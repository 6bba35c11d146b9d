// built-in system "robertson": kernel instantiations (see sde_builtin.cuh)
#include "sde_builtin.cuh"
SDE_DEFINE_BUILTIN(robertson, sde::Robertson)
