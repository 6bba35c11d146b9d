// built-in system "nonautonomous": kernel instantiations (see sde_builtin.cuh)
#include "sde_builtin.cuh"
SDE_DEFINE_BUILTIN(nonautonomous, sde::NonAutonomous)
