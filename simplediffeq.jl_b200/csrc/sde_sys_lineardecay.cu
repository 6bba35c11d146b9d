// built-in system "lineardecay": kernel instantiations (see sde_builtin.cuh)
#include "sde_builtin.cuh"
SDE_DEFINE_BUILTIN(lineardecay, sde::LinearDecay)
