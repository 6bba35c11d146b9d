// built-in system "scalargrowth": kernel instantiations (see sde_builtin.cuh)
#include "sde_builtin.cuh"
SDE_DEFINE_BUILTIN(scalargrowth, sde::ScalarGrowth)
