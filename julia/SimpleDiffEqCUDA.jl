# SimpleDiffEqCUDA.jl -- Julia host shim over libsimplediffeq_cuda (include/simplediffeq_cuda.h).
#
# STATUS: UNTESTED.  The build container has no Julia toolchain (`julia` is not installed and
# there is no network), so this file has never been executed.  It is the binding a maintainer
# of SciML/SimpleDiffEq.jl would add; INTEGRATION.md walks through it.  The Python package
# `simplediffeq.jl_b200/api.py` is the same logic in the language that could be tested here.
#
# What it does: overloads ensemble solves of the GPUSimple* algorithms so that
#     solve(EnsembleProblem(prob; prob_func), GPUSimpleTsit5(), EnsembleCUDAB200();
#           trajectories = N, dt = ..., saveat = ..., save_everystep = ...)
# evaluates prob_func on the host for i = 1:N, packs u0 / p as structure-of-arrays, makes ONE
# ccall into the C ABI (all trajectories cross the boundary together) and wraps the raw output
# into an EnsembleSolution of `build_solution` objects (retcode Default), like the reference's
# per-trajectory `solve` methods (src/tsit5/gpuatsit5.jl:136-140, :326-329) do.
module SimpleDiffEqCUDA

using SimpleDiffEq, SciMLBase, StaticArrays
import SciMLBase: __solve, EnsembleProblem, EnsembleSolution, build_solution, ReturnCode

const libsde = get(ENV, "LIBSIMPLEDIFFEQ_CUDA", "libsimplediffeq_cuda")

# ---- mirror of sde_options_t (field order and types must match the header) ---------------------
struct SdeOptions
    alg::Int32
    dtype::Int32
    save_mode::Int32
    layout::Int32
    compat::Int32
    reserved::Int32
    n_traj::Int64
    t0::Float64
    tf::Float64
    dt::Float64
    abstol::Float64
    reltol::Float64
    n_steps::Int64
    tgrid::Ptr{Cvoid}
    saveat::Ptr{Cvoid}
    n_save::Int64
    max_attempts::Int64
    out_capacity::Int64
end

const SDE_SAVE_ENDPOINT, SDE_SAVE_SAVEAT, SDE_SAVE_EVERYSTEP = Int32(0), Int32(1), Int32(2)
const SDE_LAYOUT_TRAJ_MAJOR, SDE_LAYOUT_SOA = Int32(0), Int32(1)
# `compat` keyword: 0 = the library picks the step-size controller (literal, bit-exact with a CPU run, for Float32
# states and reltol <= 1e-11; log2-domain otherwise); the flags force one of them
const SDE_COMPAT_FIX_VERN9_INTERP, SDE_COMPAT_STRICT_CONTROLLER, SDE_COMPAT_LOG2_CONTROLLER = Int32(1), Int32(2), Int32(4)
const SDE_COMPAT_FAST_RHS = Int32(8)    # contracted right-hand side: faster, no longer bit-identical to the CPU method
const SDE_COMPAT_FAST_STAGES = Int32(16)    # fixed-step Tsit5, endpoint only: step size folded into the stage coefficients (same caveat)

alg_id(::GPUSimpleTsit5) = Int32(0)
alg_id(::GPUSimpleATsit5) = Int32(1)
alg_id(::GPUSimpleRK4) = Int32(2)
alg_id(::GPUSimpleVern7) = Int32(3)
alg_id(::GPUSimpleAVern7) = Int32(4)
alg_id(::GPUSimpleVern9) = Int32(5)
alg_id(::GPUSimpleAVern9) = Int32(6)
alg_id(::GPUSimpleEuler) = Int32(7)
const SavesEveryStep = Union{GPUSimpleRK4, GPUSimpleEuler}     # no saveat / save_everystep keywords in the reference
is_adaptive(alg) = alg isa Union{GPUSimpleATsit5, GPUSimpleAVern7, GPUSimpleAVern9}

"Ensemble algorithm selecting the B200 library; `devices` = CUDA ordinals to shard over."
struct EnsembleCUDAB200 <: SciMLBase.EnsembleAlgorithm
    devices::Vector{Cint}
end
EnsembleCUDAB200() = EnsembleCUDAB200(Cint[0])

last_error() = unsafe_string(ccall((:sde_last_error, libsde), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("libsimplediffeq_cuda: " * last_error())

# ---- right-hand sides ---------------------------------------------------------------------------
"A device right-hand side: built-in registry entry or CUDA-C source (NVRTC)."
struct DeviceRHS
    handle::Ptr{Cvoid}
    n_state::Int
    n_param::Int
end

function builtin_rhs(name::AbstractString)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sde_system_builtin, libsde), Cint, (Cstring, Ref{Ptr{Cvoid}}), name, h))
    ns, np = Ref{Cint}(0), Ref{Cint}(0)
    check(ccall((:sde_system_dims, libsde), Cint, (Ptr{Cvoid}, Ref{Cint}, Ref{Cint}), h[], ns, np))
    DeviceRHS(h[], ns[], np[])
end

"`src` defines `__device__ void rhs(real* du, const real* u, const real* p, real t)`."
function cuda_rhs(src::AbstractString, n_state::Integer, n_param::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    log = Vector{UInt8}(undef, 16384)
    rc = ccall((:sde_system_nvrtc, libsde), Cint,
        (Cstring, Cint, Cint, Ref{Ptr{Cvoid}}, Ptr{UInt8}, Csize_t),
        src, n_state, n_param, h, log, length(log))
    rc == 0 || error("NVRTC: " * last_error())
    DeviceRHS(h[], n_state, n_param)
end

# The ODEProblem's `f` stays an ordinary Julia function (used by the CPU path and by prob_func);
# the device RHS that implements the same formula is attached through this registry.
const DEVICE_RHS = IdDict{Any, DeviceRHS}()
register_device_rhs!(f, rhs::DeviceRHS) = (DEVICE_RHS[f] = rhs)

# ---- zero-copy views of series outputs ------------------------------------------------------------
# SDE_LAYOUT_SOA (out[i, c, slot] in column-major Julia = out_u[slot][c][i] in C) is the layout that
# reaches 90 % of HBM bandwidth (DESIGN.md section 3); this wrapper presents trajectory i of such an
# array as the `Vector{SVector{N,T}}`-like object `sol.u` is expected to be, without copying.
struct SoASeries{N, T} <: AbstractVector{SVector{N, T}}
    data::Array{T, 3}        # (n_traj, N, n_slots)
    i::Int
    len::Int
end
Base.size(s::SoASeries) = (s.len,)
Base.@propagate_inbounds Base.getindex(s::SoASeries{N, T}, j::Int) where {N, T} =
    SVector{N, T}(ntuple(c -> s.data[s.i, c, j], N))

# ---- the ensemble solve ---------------------------------------------------------------------------
function SciMLBase.__solve(ensembleprob::EnsembleProblem, alg::Union{GPUSimpleTsit5, GPUSimpleATsit5,
            GPUSimpleRK4, GPUSimpleEuler, GPUSimpleVern7, GPUSimpleAVern7, GPUSimpleVern9, GPUSimpleAVern9},
        ensemblealg::EnsembleCUDAB200;
        trajectories, dt = alg isa SavesEveryStep ? error("dt is required for this algorithm") : 0.1f0,
        abstol = 1.0f-6, reltol = 1.0f-3, saveat = nothing, save_everystep = true,
        layout = SDE_LAYOUT_TRAJ_MAJOR, compat = 0, kwargs...)
    prob = ensembleprob.prob
    @assert !SciMLBase.isinplace(prob)
    T = eltype(prob.u0)
    T in (Float64, Float32) || error("only Float64 / Float32 states run on the device; use the reference's own method")
    f = prob.f isa SciMLBase.ODEFunction ? prob.f.f : prob.f
    rhs = get(DEVICE_RHS, f, nothing)
    rhs === nothing && error("no device RHS registered for this f (register_device_rhs!)")
    N, NP, n = rhs.n_state, rhs.n_param, Int(trajectories)

    # prob_func on the host, i = 1:N; only u0 and p may change
    u0 = Matrix{T}(undef, n, N)          # column-major: u0[i, c] == SoA [c][i]
    p = Matrix{T}(undef, n, NP)
    for i in 1:n
        pi = ensembleprob.prob_func(prob, i, 1)
        pi.tspan == prob.tspan || error("prob_func may only change u0 and p on the device path")
        u0[i, :] .= pi.u0
        NP > 0 && (p[i, :] .= pi.p)
    end

    t0, tf = T(prob.tspan[1]), T(prob.tspan[2])
    dtT = T(dt)
    adaptive = is_adaptive(alg)
    save_mode = alg isa SavesEveryStep ? SDE_SAVE_EVERYSTEP :
                saveat !== nothing ? SDE_SAVE_SAVEAT :
                save_everystep ? SDE_SAVE_EVERYSTEP : SDE_SAVE_ENDPOINT
    tgrid = adaptive ? T[] : collect(T, t0:dtT:tf)          # _ts = tspan[1]:dt:tspan[2]
    sa = saveat === nothing ? T[] : collect(T, saveat)
    n_steps = adaptive ? 0 : length(tgrid) - 1
    nacc, nrej, ret = zeros(Int32, n), zeros(Int32, n), zeros(Int32, n)

    function call(mode, capacity, out_u, out_t)
        GC.@preserve tgrid sa begin
            opt = Ref(SdeOptions(alg_id(alg), T === Float64 ? 0 : 1, mode, layout, compat, 0, n,
                t0, tf, dtT, T(abstol), T(reltol), n_steps,
                isempty(tgrid) ? C_NULL : pointer(tgrid), isempty(sa) ? C_NULL : pointer(sa), length(sa), 0, capacity))
            check(ccall((:sde_solve, libsde), Cint,
                (Ptr{Cvoid}, Ref{SdeOptions}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32},
                    Ptr{Int32}, Ptr{Cint}, Cint),
                rhs.handle, opt, u0, p, out_u, out_t, nacc, nrej, ret, ensemblealg.devices, length(ensemblealg.devices)))
            any(==(1), ret) && error("dt<dtmin")            # what the reference throws (gpuatsit5.jl:256)
            ts_fixed = T[]
            if !adaptive
                ts_fixed = Vector{T}(undef, max(n_steps + 1, length(sa), 2))
                nw = Ref{Int64}(0)
                check(ccall((:sde_fixed_times, libsde), Cint, (Ref{SdeOptions}, Ptr{Cvoid}, Int64, Ref{Int64}),
                    opt, ts_fixed, length(ts_fixed), nw))
                resize!(ts_fixed, nw[])
            end
            return ts_fixed
        end
    end

    # adaptive save_everystep = true pushes every accepted step (gpuatsit5.jl:301-303): the step sequence is
    # deterministic, so an endpoint-only pass gives the exact output sizes
    capacity = 0
    if adaptive && save_mode == SDE_SAVE_EVERYSTEP
        call(SDE_SAVE_ENDPOINT, 0, Matrix{T}(undef, n, N), Vector{T}(undef, n))
        capacity = Int(maximum(nacc)) + 1
    end
    slots = save_mode == SDE_SAVE_SAVEAT ? length(sa) :
            save_mode == SDE_SAVE_EVERYSTEP ? (adaptive ? capacity : n_steps + 1) : 1
    t_series = adaptive && save_mode == SDE_SAVE_EVERYSTEP
    soa = layout == SDE_LAYOUT_SOA
    out_u = save_mode == SDE_SAVE_ENDPOINT ? Matrix{T}(undef, n, N) :
            soa ? Array{T}(undef, n, N, slots) : Array{T}(undef, N, slots, n)
    out_t = t_series ? (soa ? Matrix{T}(undef, n, slots) : Matrix{T}(undef, slots, n)) : Vector{T}(undef, n)
    ts_fixed = call(save_mode, capacity, out_u, out_t)

    SV = SVector{N, T}
    sols = map(1:n) do i
        pi = ensembleprob.prob_func(prob, i, 1)
        if save_mode == SDE_SAVE_ENDPOINT
            us = [SV(pi.u0), SV(ntuple(c -> out_u[i, c], N))]
            ts = adaptive ? T[t0, out_t[i]] : ts_fixed
        else
            len = t_series ? Int(nacc[i]) + 1 : slots      # adaptive every-step: naccept + 1 states were pushed
            us = soa ? SoASeries{N, T}(out_u, i, len) : collect(reinterpret(SV, vec(view(out_u, :, :, i))))[1:len]
            ts = save_mode == SDE_SAVE_SAVEAT ? sa : t_series ? (soa ? out_t[i, 1:len] : out_t[1:len, i]) : ts_fixed
        end
        build_solution(pi, alg, ts, us; calculate_error = false)   # retcode stays ReturnCode.Default
    end
    return EnsembleSolution(sols, 0.0, true)
end

# ==================================================================================================
# SimpleEM: solve(EnsembleProblem(SDEProblem(f, g, u0, tspan, p); prob_func), SimpleEM(), EnsembleCUDAB200();
#                 trajectories, dt, seed = rand(UInt64))      (replaces src/euler_maruyama.jl:48-94 per trajectory;
#                 a fresh Philox key per solve like the reference's randn, an explicit seed reproduces a run)
# ==================================================================================================
struct SdeEmOptions
    dtype::Int32
    save_mode::Int32
    layout::Int32
    noise_mode::Int32
    n_traj::Int64
    t0::Float64
    dt::Float64
    n_steps::Int64
    seed::UInt64
    traj_offset::Int64
end

"Device drift + diffusion: built-in SDE registry entry or CUDA-C source defining `rhs` and `noise`."
struct DeviceSDE
    handle::Ptr{Cvoid}
    n_state::Int
    n_param::Int
    n_noise::Int
end

function builtin_sde(name::AbstractString)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sde_em_system_builtin, libsde), Cint, (Cstring, Ref{Ptr{Cvoid}}), name, h))
    a, b, c, d = Ref{Cint}(0), Ref{Cint}(0), Ref{Cint}(0), Ref{Cint}(0)
    check(ccall((:sde_em_system_dims, libsde), Cint, (Ptr{Cvoid}, Ref{Cint}, Ref{Cint}, Ref{Cint}, Ref{Cint}), h[], a, b, c, d))
    DeviceSDE(h[], a[], b[], c[])
end

function cuda_sde(src::AbstractString, n_state::Integer, n_param::Integer; n_noise::Integer = n_state, diagonal::Bool = true)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    log = Vector{UInt8}(undef, 16384)
    rc = ccall((:sde_em_system_nvrtc, libsde), Cint,
        (Cstring, Cint, Cint, Cint, Cint, Ref{Ptr{Cvoid}}, Ptr{UInt8}, Csize_t),
        src, n_state, n_param, n_noise, diagonal ? 1 : 0, h, log, length(log))
    rc == 0 || error("NVRTC: " * last_error())
    DeviceSDE(h[], n_state, n_param, n_noise)
end

const DEVICE_SDE = IdDict{Any, DeviceSDE}()
register_device_sde!(f, sde::DeviceSDE) = (DEVICE_SDE[f] = sde)

function SciMLBase.__solve(ensembleprob::EnsembleProblem, alg::SimpleEM, ensemblealg::EnsembleCUDAB200;
        trajectories, dt = error("dt required for SimpleEM"), seed::Integer = rand(UInt64), kwargs...)
    prob = ensembleprob.prob
    @assert !SciMLBase.isinplace(prob)
    T = eltype(prob.u0)
    T in (Float64, Float32) || error("only Float64 / Float32 states run on the device")
    f = prob.f isa SciMLBase.SDEFunction ? prob.f.f : prob.f
    sde = get(DEVICE_SDE, f, nothing)
    sde === nothing && error("no device SDE registered for this f (register_device_sde!)")
    N, NP, n = sde.n_state, sde.n_param, Int(trajectories)
    u0 = Matrix{T}(undef, n, N)
    p = Matrix{T}(undef, n, NP)
    for i in 1:n
        pi = ensembleprob.prob_func(prob, i, 1)
        pi.tspan == prob.tspan || error("prob_func may only change u0 and p on the device path")
        u0[i, :] .= pi.u0
        NP > 0 && (p[i, :] .= pi.p)
    end
    t0, dtT = T(prob.tspan[1]), T(dt)
    nst = Int((T(prob.tspan[2]) - t0) / dtT)                 # n - 1; InexactError as in the reference (:66)
    out_u = Array{T}(undef, N, nst + 1, n)                    # trajectory major, every state (:67)
    opt = Ref(SdeEmOptions(T === Float64 ? 0 : 1, SDE_SAVE_EVERYSTEP, SDE_LAYOUT_TRAJ_MAJOR, 0, n, t0, dtT, nst,
        seed % UInt64, 0))
    check(ccall((:sde_em_solve, libsde), Cint,
        (Ptr{Cvoid}, Ref{SdeEmOptions}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cint}, Cint),
        sde.handle, opt, u0, p, C_NULL, out_u, ensemblealg.devices, length(ensemblealg.devices)))
    ts = [muladd(T(i), dtT, t0) for i in 0:nst]               # :68 under @muladd
    sols = map(1:n) do i
        pi = ensembleprob.prob_func(prob, i, 1)
        us = prob.u0 isa Number ? vec(out_u[1, :, i]) : collect(reinterpret(SVector{N, T}, vec(view(out_u, :, :, i))))
        build_solution(pi, alg, ts, us; calculate_error = false)
    end
    return EnsembleSolution(sols, 0.0, true)
end

end # module
