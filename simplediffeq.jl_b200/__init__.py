"""B200-native ensemble ODE integrator: drop-in for the GPUSimple* solver family of
SciML/SimpleDiffEq.jl (GPUSimpleTsit5/ATsit5, GPUSimpleRK4, GPUSimpleVern7/AVern7,
GPUSimpleVern9/AVern9).  Hot path: hand-written sm_100a kernels behind the C ABI
include/simplediffeq_cuda.h (libsimplediffeq_cuda.so, built in-tree by build.py).
Import as `simplediffeq_b200` (see simplediffeq_b200.py at the repository root)."""
from . import _lib
from .api import (CudaRHS, EnsembleProblem, EnsembleSolution, GPUSimpleATsit5, GPUSimpleAVern7,
                  GPUSimpleAVern9, GPUSimpleEuler, GPUSimpleRK4, GPUSimpleTsit5, GPUSimpleVern7, GPUSimpleVern9,
                  ODEProblem, ODESolution, System, builtin_system, remake, solve, solve_arrays,
                  solve_device, systems)
from .em import (CudaSDE, EMEnsembleSolution, SDEProblem, SDESystem, SimpleEM, em_noise, em_steps, em_times,
                 sde_systems, solve_em, solve_em_arrays, solve_em_device)
from .jlrange import JuliaRange, jl_range

__all__ = ["CudaRHS", "EnsembleProblem", "EnsembleSolution", "GPUSimpleATsit5", "GPUSimpleAVern7",
           "GPUSimpleAVern9", "GPUSimpleEuler", "GPUSimpleRK4", "GPUSimpleTsit5", "GPUSimpleVern7", "GPUSimpleVern9",
           "ODEProblem", "ODESolution", "System", "builtin_system", "remake", "solve", "solve_arrays",
           "solve_device", "systems", "JuliaRange", "jl_range",
           "CudaSDE", "EMEnsembleSolution", "SDEProblem", "SDESystem", "SimpleEM", "em_noise", "em_steps", "em_times",
           "sde_systems", "solve_em", "solve_em_arrays", "solve_em_device"]
