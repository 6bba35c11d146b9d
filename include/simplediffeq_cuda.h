/* simplediffeq_cuda.h -- C ABI of libsimplediffeq_cuda (B200 / sm_100a ensemble ODE integrator).
 *
 * Drop-in boundary for the GPU solver family of SciML/SimpleDiffEq.jl v1.16.3.  One call solves a
 * whole ensemble; each entry point below names the reference interface it replaces
 * (paths relative to the reference repository):
 *
 *   sde_solve / sde_solve_device
 *       DiffEqBase.solve(prob::ODEProblem, alg::GPUSimpleTsit5;  saveat, save_everystep, dt)                 src/tsit5/gpuatsit5.jl:55-147
 *       DiffEqBase.solve(prob::ODEProblem, alg::GPUSimpleATsit5; dt, saveat, save_everystep, abstol, reltol) src/tsit5/gpuatsit5.jl:205-336
 *       DiffEqBase.solve(prob::ODEProblem, alg::GPUSimpleRK4;    dt)                                         src/rk4/gpurk4.jl:53-98
 *       DiffEqBase.solve(prob::ODEProblem, alg::GPUSimpleEuler;  dt)                                         src/euler/gpueuler.jl:53-90
 *       DiffEqBase.solve(prob::ODEProblem, alg::GPUSimpleVern7 / GPUSimpleAVern7; ...)                       src/verner/gpuvern7.jl:55-242, :300-536
 *       DiffEqBase.solve(prob::ODEProblem, alg::GPUSimpleVern9 / GPUSimpleAVern9; ...)                       src/verner/gpuvern9.jl:55-353, :411-779
 *     called once per trajectory by SciMLBase's ensemble driver (batch_func) in the reference; here
 *     ALL trajectories of `solve(EnsembleProblem(prob; prob_func), alg; trajectories, ...)` cross the
 *     boundary in one call.
 *   sde_system_builtin / sde_system_nvrtc
 *       `prob.f` (an arbitrary Julia callable f(u,p,t) in the reference, src/tsit5/gpuatsit5.jl:65):
 *       a built-in registry entry or a CUDA-C `__device__` function JIT-compiled with NVRTC.
 *   per-trajectory retcode
 *       the reference's exceptions: error("dt<dtmin") src/tsit5/gpuatsit5.jl:256,
 *       src/verner/gpuvern7.jl:358, src/verner/gpuvern9.jl:466; ReturnCode.Default otherwise
 *       (test/gpu_ode_regression.jl:24-25).
 *   sde_em_solve / sde_em_solve_device / sde_em_system_* / sde_em_noise*   (second half of this file)
 *       DiffEqBase.solve(prob::SDEProblem{uType,tType,false}, alg::SimpleEM; dt)   src/euler_maruyama.jl:48-94
 *
 * All functions return SDE_OK (0) or a negative error code; sde_last_error() gives the message of
 * the calling thread's last failure.  Nothing throws across this boundary.  The library is
 * re-entrant; system handles may be shared between threads.
 */
#ifndef SIMPLEDIFFEQ_CUDA_H
#define SIMPLEDIFFEQ_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDE_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define SDE_API __attribute__((visibility("default")))
#else
#define SDE_API
#endif

/* status codes */
enum {
  SDE_OK = 0,
  SDE_ERR_INVALID = -1,     /* bad argument */
  SDE_ERR_CUDA = -2,        /* CUDA runtime error (no device, launch failure, ...) */
  SDE_ERR_NVRTC = -3,       /* user RHS failed to compile */
  SDE_ERR_UNSUPPORTED = -4, /* combination not provided by this build */
  SDE_ERR_NOMEM = -5
};

/* algorithms (the reference's GPUSimple* singletons) */
enum {
  SDE_ALG_TSIT5 = 0,  /* GPUSimpleTsit5   fixed step  */
  SDE_ALG_ATSIT5 = 1, /* GPUSimpleATsit5  adaptive    */
  SDE_ALG_RK4 = 2,    /* GPUSimpleRK4     fixed step, always saves every step in the reference */
  SDE_ALG_VERN7 = 3,  /* GPUSimpleVern7   fixed step  */
  SDE_ALG_AVERN7 = 4, /* GPUSimpleAVern7  adaptive    */
  SDE_ALG_VERN9 = 5,  /* GPUSimpleVern9   fixed step  */
  SDE_ALG_AVERN9 = 6, /* GPUSimpleAVern9  adaptive    */
  SDE_ALG_EULER = 7   /* GPUSimpleEuler   fixed step, always saves every step (src/euler/gpueuler.jl:53-90) */
};

enum { SDE_F64 = 0, SDE_F32 = 1 };

/* what is saved */
enum {
  SDE_SAVE_ENDPOINT = 0,  /* saveat === nothing && save_everystep = false : final state only */
  SDE_SAVE_SAVEAT = 1,    /* saveat = [...] : dense output at the given times */
  SDE_SAVE_EVERYSTEP = 2  /* saveat === nothing && save_everystep = true.  Fixed step: n_steps + 1 states,
                             slot 0 = u0.  Adaptive: slot 0 = u0, slot k = state after the k-th accepted step,
                             at most opt->out_capacity slots per trajectory (see sde_solve) */
};

/* layout of series outputs (SAVEAT / EVERYSTEP); n_out = slots per trajectory */
enum {
  SDE_LAYOUT_TRAJ_MAJOR = 0, /* out_u[(i*n_out + s)*n_state + c] : per-trajectory Vector{SVector} */
  SDE_LAYOUT_SOA = 1         /* out_u[(s*n_state + c)*n_traj + i] : coalesced, trajectory fastest */
};

/* per-trajectory return codes */
enum {
  SDE_RET_DEFAULT = 0,  /* ReturnCode.Default */
  SDE_RET_DTMIN = 1,    /* the reference would throw error("dt<dtmin") */
  SDE_RET_MAXITERS = 2, /* max_attempts exhausted (no counterpart in the reference) */
  SDE_RET_OUTPUT_FULL = 3 /* adaptive SDE_SAVE_EVERYSTEP: the trajectory took more than out_capacity - 1 accepted
                             steps; the first out_capacity states are stored, naccept tells how many are needed */
};

/* compat flags: 0 = reproduce the reference exactly, including its quirks.
   Step-size controller of the adaptive algorithms (src/tsit5/gpuatsit5.jl:276-292): two implementations.
     literal : the reference's arithmetic operation for operation -- IEEE divisions, sqrt, `EEst^beta1`,
               `qold^beta2` -- with the pow / powf of the libm a CPU run on an x86-64 Linux host uses (the
               table-driven pow of glibc >= 2.28 = ARM optimized-routines, FMA variant; one log half shared
               by both powers).  Reproduces the CPU oracle bit for bit: states, times, accepted and rejected
               counts.  ~1.3x (AVern9) to 1.5x (ATsit5) the device time of the log2 one.
     log2    : the same formulas evaluated in the log2 domain (1 log2 + 1 exp2, no division, no sqrt):
               ~1e-15 relative difference in the next dt; accept / reject decisions unchanged except within
               ~1e-15 of EEst = 1.
   With neither flag the library chooses: the literal controller where the step sequence hangs on the last
   bit of the powers and the log2 one cannot reproduce step counts -- Float32 states, or reltol <= 1e-11
   (BASELINE config 4, AVern9 at 1e-12) -- and the log2 controller everywhere else (identical accepted-step
   counts on the BASELINE sweeps at 1e-6 ... 1e-10).  DESIGN.md sections 2 and 6. */
enum {
  SDE_COMPAT_FIX_VERN9_INTERP = 1, /* fixed-step GPUSimpleVern9 + saveat: use stages 8..15 in the dense
                                      output (the reference uses k2..k9, src/verner/gpuvern9.jl:216-331) */
  SDE_COMPAT_STRICT_CONTROLLER = 2, /* adaptive: force the literal controller */
  SDE_COMPAT_LOG2_CONTROLLER = 4,   /* adaptive: force the log2-domain controller (exclusive with the above) */
  SDE_COMPAT_FAST_RHS = 8,          /* throughput beyond the reference-exact ceiling: the right-hand side f may be
                                       contracted into fused multiply-adds (the reference never applies @muladd to
                                       f).  Built-in lorenz: 8 -> 6 FP64 instructions per evaluation, a fixed-step
                                       Tsit5 step 126 -> 114 (~ +10 % steps/s); vanderpol 5 -> 3; user CUDA-C
                                       systems are compiled with --fmad=true; other built-ins: no effect.  Results
                                       are no longer bit-identical to the reference's: on BASELINE config 2's
                                       sweep <= 1e-12 relative (median 6e-16) except within 0.01 of the homoclinic
                                       bifurcation at rho = 13.926 (max 6e-12); unbounded for chaotic trajectories,
                                       like any rounding change.  Off by default. */
  SDE_COMPAT_FAST_STAGES = 16       /* throughput beyond the reference-exact ceiling, second step: fixed-step
                                       GPUSimpleTsit5 (every save mode) folds the step size into the stage
                                       coefficients, tmp = uprev + sum_j (dt a_ij) k_j instead of the reference's
                                       uprev + dt (sum_j a_ij k_j): 21 N instead of 26 N + 1 FP64 instructions per
                                       step for the stage sums, any system (built-in or CUDA-C).  With
                                       SDE_COMPAT_FAST_RHS on lorenz: 126 -> 99 per step, 1.84e11 steps/s on BASELINE
                                       config 2 (+25 % over the reference-exact kernel).  Every term is then rounded
                                       at the magnitude of the state, so the deviation from the reference is that of
                                       one more rounding of u per stage: on BASELINE config 2's sweep median 5e-15
                                       relative, 99.9 % of the trajectories <= 6e-13, <= 1e-12 except within 0.025
                                       of rho = 13.926 (max 2.3e-11).  Dense output (saveat) is the reference's
                                       formula on the deviating stages.  Helps FP64-bound launches; an HBM-bound
                                       saveat launch gets a few per cent slower (more registers).  Other algorithms
                                       ignore the flag.  Off by default. */
};

typedef struct sde_system_s* sde_system_t;

typedef struct sde_options {
  int32_t alg;        /* SDE_ALG_* */
  int32_t dtype;      /* SDE_F64 / SDE_F32: element type of every buffer and of all arithmetic */
  int32_t save_mode;  /* SDE_SAVE_* */
  int32_t layout;     /* SDE_LAYOUT_* (series outputs) */
  int32_t compat;     /* SDE_COMPAT_* flags */
  int32_t reserved;
  int64_t n_traj;     /* `trajectories` */
  double t0, tf;      /* prob.tspan (converted to dtype) */
  double dt;          /* dt (fixed step) / initial dt (adaptive); reference default 0.1f0 */
  double abstol;      /* adaptive; reference default 1f-6 */
  double reltol;      /* adaptive; reference default 1f-3 */
  int64_t n_steps;    /* fixed step: length(t0:dt:tf) - 1 */
  const void* tgrid;  /* fixed step: HOST array, the n_steps+1 elements of t0:dt:tf in dtype.
                         May be NULL: then t0 + k*dt (rounded product, rounded sum) is used. */
  const void* saveat; /* SDE_SAVE_SAVEAT: HOST array of n_save times in dtype */
  int64_t n_save;
  int64_t max_attempts; /* adaptive: attempts (accepted + rejected) per trajectory; 0 = unlimited like the
                           reference (the int32 range of the naccept / nreject counters) */
  int64_t out_capacity; /* adaptive SDE_SAVE_EVERYSTEP: slots per trajectory in out_u / out_t (>= 1) */
} sde_options_t;

/* ---- library / device ------------------------------------------------------------------- */
SDE_API int sde_version(void);
SDE_API const char* sde_last_error(void);
SDE_API int sde_device_count(int* count);

/* ---- right-hand sides --------------------------------------------------------------------- */
/* names: "lorenz", "vanderpol", "robertson", "nbody", "lineardecay", "scalargrowth",
 * "nonautonomous".  Handles of built-ins are static; sde_system_free on them is a no-op. */
SDE_API int sde_system_builtin(const char* name, sde_system_t* out);

/* User RHS.  `src` is CUDA C++ defining
 *     __device__ void rhs(real* du, const real* u, const real* p, real t)
 * (`real` is typedef'd to double or float by the library).  It is compiled for sm_100a with
 * --fmad=false so that the user's arithmetic rounds as written (the reference does not fuse f).
 * Kernels are compiled lazily per (algorithm, dtype, save mode) and cached in the handle.
 * `log`/`log_len`: optional buffer receiving the NVRTC log of a failed compile of the syntax
 * check done here. */
SDE_API int sde_system_nvrtc(const char* src, int n_state, int n_param, sde_system_t* out, char* log,
                     size_t log_len);
SDE_API int sde_system_dims(sde_system_t sys, int* n_state, int* n_param);
SDE_API void sde_system_free(sde_system_t sys);
/* Compile (without launching) the kernel a solve with these options would use.  Works without a
 * GPU for NVRTC systems; for built-ins it only checks that the kernel exists in this build. */
SDE_API int sde_system_prepare(sde_system_t sys, const sde_options_t* opt);

/* ---- solve, host buffers ------------------------------------------------------------------ */
/* u0: [n_state][n_traj] (SoA), p: [n_param][n_traj] (SoA), HOST memory (pinned or pageable).
 * out_u: ENDPOINT -> [n_state][n_traj] (SoA final states)
 *        SAVEAT   -> n_out = n_save slots per trajectory, layout per opt->layout
 *        EVERYSTEP-> fixed step: n_out = n_steps + 1 slots per trajectory; adaptive: n_out = out_capacity
 *                    slots, of which min(naccept + 1, n_out) are written (run SDE_SAVE_ENDPOINT first to
 *                    learn naccept when no bound is known: the step sequence is deterministic)
 *        series slots that the reference would leave `undef` are NaN.
 * out_t: adaptive ENDPOINT / SAVEAT: [n_traj] final time of each trajectory (== tf unless retcode != 0);
 *        adaptive EVERYSTEP: the time of every stored slot, same layout as out_u without the component
 *        axis ([n_traj][n_out] or [n_out][n_traj]); may be NULL.
 *        fixed step: ignored (times are trajectory independent: see sde_fixed_times).
 * naccept/nreject/retcode: [n_traj] int32, each may be NULL; fixed-step algorithms write zeros
 *        (no step control, retcode Default).
 * devices/n_dev: CUDA device ordinals to shard over, one host thread per device, no collective;
 *        NULL/0 = current device only.  Fixed-step algorithms: contiguous index ranges
 *        [g*N/G, (g+1)*N/G).  Adaptive algorithms: contiguous pieces handed out from a shared cursor
 *        (step counts are not uniform along a sweep).  Results do not depend on the assignment. */
SDE_API int sde_solve(sde_system_t sys, const sde_options_t* opt, const void* u0, const void* p,
              void* out_u, void* out_t, int32_t* naccept, int32_t* nreject, int32_t* retcode,
              const int* devices, int n_dev);

/* ---- solve, device-resident buffers ------------------------------------------------------- */
/* Same contract with DEVICE pointers on the current device; opt->tgrid / opt->saveat remain HOST
 * arrays (small; uploaded and cached per call).  `ld_in` / `ld_out` are the component strides of the
 * SoA inputs / endpoint (or SoA series) outputs, so that a shard of a larger array can be solved
 * in place (pass n_traj = shard length, pointers offset to the shard start).  Work is enqueued
 * on `stream` (a cudaStream_t, NULL = default stream) and the call returns without
 * synchronising when `async` != 0. */
SDE_API int sde_solve_device(sde_system_t sys, const sde_options_t* opt, const void* d_u0, const void* d_p,
                     int64_t ld_in, void* d_out_u, int64_t ld_out, void* d_out_t,
                     int32_t* d_naccept, int32_t* d_nreject, int32_t* d_retcode, void* stream,
                     int async);

/* Times the reference would put in sol.t for a fixed-step solve (identical for every trajectory):
 * ENDPOINT -> 2 values [t0, t_end]; EVERYSTEP -> n_steps+1 values; SAVEAT -> copy of saveat.
 * `out` is a HOST array in dtype with room for `n` values; returns the count written in *n_written. */
SDE_API int sde_fixed_times(const sde_options_t* opt, void* out, int64_t n, int64_t* n_written);

/* Pinned host memory helpers (so that the H2D/D2H copies of sde_solve run at full PCIe rate). */
SDE_API int sde_host_alloc(void** ptr, size_t bytes);
SDE_API int sde_host_free(void* ptr);

/* Measured FMA-pipe peak of the current device (dense unrolled DFMA/FFMA chains, best of 3 after
 * warm-up), in TFLOP/s; *ms = duration of the best run.  The roofline denominator bench.py reports. */
SDE_API int sde_probe_fma_peak(int dtype, double* tflops, double* ms);

/* Number of kernels launched by this process through the library (all threads). */
SDE_API int64_t sde_launch_count(void);

/* Device buffers of sde_solve come from one stream-ordered memory pool per device, so that repeated
 * solves do not pay cudaMalloc/cudaFree (the reference allocates per trajectory through Julia's GC;
 * there is no counterpart to cite).  After each solve the pool keeps at most SDE_POOL_KEEP_MB
 * (environment, default 4096) megabytes; sde_trim() returns everything to the driver. */
SDE_API int sde_trim(void);

/* ============================================================================================
 * SimpleEM: fixed-step Euler-Maruyama ensembles for SDEs du = f(u,p,t) dt + g(u,p,t) dW.
 * Replaces `DiffEqBase.solve(prob::SDEProblem{uType,tType,false}, alg::SimpleEM; dt)`
 * (src/euler_maruyama.jl:48-94, the out-of-place method), called once per trajectory by the
 * reference's ensemble driver.  n_steps = Int((tspan[2]-tspan[1])/dt) (:66) is computed by the host
 * layer (Julia raises InexactError when the quotient is not an integer); the states are
 * t_i = muladd(i, dt, t0) (:68).  The reference keeps every state (:67): SDE_SAVE_EVERYSTEP gives
 * n_steps + 1 slots per trajectory (slot 0 = u0), SDE_SAVE_ENDPOINT only the last one.
 *
 * Noise: the reference calls randn() on Julia's task-local RNG (:77,:80,:84), which cannot be
 * reproduced outside that process.  SDE_NOISE_PHILOX draws the increments from Philox4x32-10 keyed
 * by `seed` and counted by (traj_offset + trajectory index, step, component) -- see
 * csrc/device/sde_em.cuh for the exact layout -- so results do not depend on devices, pieces or
 * launch geometry; SDE_NOISE_PROVIDED reads standard normals from the caller:
 * noise[(step * n_noise + m)][trajectory] (SoA).  sde_em_noise() returns the normals a PHILOX solve
 * with the same options consumes, in that layout.
 * ============================================================================================ */
typedef struct sde_em_system_s* sde_em_system_t;

#define SDE_NOISE_PHILOX 0
#define SDE_NOISE_PROVIDED 1

typedef struct sde_em_options {
  int32_t dtype;        /* SDE_F64 | SDE_F32 */
  int32_t save_mode;    /* SDE_SAVE_EVERYSTEP (the reference) | SDE_SAVE_ENDPOINT */
  int32_t layout;       /* SDE_SAVE_EVERYSTEP output: SDE_LAYOUT_TRAJ_MAJOR | SDE_LAYOUT_SOA */
  int32_t noise_mode;   /* SDE_NOISE_PHILOX | SDE_NOISE_PROVIDED */
  int64_t n_traj;
  double t0, dt;
  int64_t n_steps;
  uint64_t seed;        /* PHILOX key */
  int64_t traj_offset;  /* global index of trajectory 0 (PHILOX counter); 0 for a whole ensemble */
} sde_em_options_t;

/* Built-in SDE systems: "gbm" (f = p1*u, g = p2*u; the docstring example src/euler_maruyama.jl:27-28),
 * "linadd1" / "linadd2" (f = p1*u, g = p2; test/simpleem_tests.jl:4-5,16, scalar / 2 components),
 * "ou" (f = p1*(p2-u), g = p3), "nondiag2x4" (f = p1.*u, G = the 2x4 matrix of test/simpleem_tests.jl:33-47). */
SDE_API int sde_em_system_builtin(const char* name, sde_em_system_t* out);

/* User SDE: CUDA-C source defining
 *     __device__ void rhs  (real* f, const real* u, const real* p, real t);   // drift, n_state values
 *     __device__ void noise(real* g, const real* u, const real* p, real t);   // diagonal: n_state values;
 *                                                                              // else n_state x n_noise, row major
 * compiled by NVRTC for sm_100a with --fmad=false (replaces prob.f / prob.g). */
SDE_API int sde_em_system_nvrtc(const char* src, int n_state, int n_param, int n_noise, int diagonal,
                                sde_em_system_t* out, char* log, size_t log_len);
SDE_API int sde_em_system_dims(sde_em_system_t sys, int* n_state, int* n_param, int* n_noise, int* diagonal);
SDE_API void sde_em_system_free(sde_em_system_t sys);
SDE_API int sde_em_system_prepare(sde_em_system_t sys, const sde_em_options_t* opt);

/* Host buffers: u0 [n_state][n_traj], p [n_param][n_traj] (SoA); noise (PROVIDED only, else NULL)
 * [n_steps * n_noise][n_traj]; out_u: ENDPOINT [n_state][n_traj], EVERYSTEP n_steps+1 slots per
 * trajectory in opt->layout.  devices/n_dev as in sde_solve (contiguous ranges, no collective). */
SDE_API int sde_em_solve(sde_em_system_t sys, const sde_em_options_t* opt, const void* u0, const void* p,
                         const void* noise, void* out_u, const int* devices, int n_dev);

/* Device-resident buffers on the current device (ld_* = component strides, as in sde_solve_device). */
SDE_API int sde_em_solve_device(sde_em_system_t sys, const sde_em_options_t* opt, const void* d_u0,
                                const void* d_p, int64_t ld_in, const void* d_noise, int64_t noise_ld,
                                void* d_out_u, int64_t ld_out, void* stream, int async);

/* The standard normals of the PHILOX stream: out[(step * n_noise + m)][trajectory] for the opt's
 * (dtype, seed, traj_offset, n_traj, n_steps); host array / device array with row stride ld. */
SDE_API int sde_em_noise(const sde_em_options_t* opt, int n_noise, void* out);
SDE_API int sde_em_noise_device(const sde_em_options_t* opt, int n_noise, void* d_out, int64_t ld, void* stream,
                                int async);

#ifdef __cplusplus
}
#endif
#endif /* SIMPLEDIFFEQ_CUDA_H */
