// built-in system "lorenz": kernel instantiations (see sde_builtin.cuh)
#include "sde_builtin.cuh"
SDE_DEFINE_BUILTIN(lorenz, sde::Lorenz)
