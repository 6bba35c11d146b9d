// sde_kernels.cuh -- one-trajectory-per-thread ensemble integrator bodies for sm_100a.
//
// Two kernel shapes:
//   fixed_body     GPUSimpleTsit5 / GPUSimpleRK4 / GPUSimpleVern7 / GPUSimpleVern9
//                  (every trajectory takes the same n_steps -> plain grid, no divergence)
//   adaptive_body  GPUSimpleATsit5 / GPUSimpleAVern7 / GPUSimpleAVern9
//                  persistent CTAs; every LANE pulls its next trajectory from a global atomic
//                  work queue (warp-aggregated atomicAdd) as soon as its current one reaches tf, so
//                  trajectories with unequal step counts rebalance; the warp leaves when
//                  __all_sync says no lane has work and the queue is drained.
//
// State, parameters and all stage vectors live in registers (arrays indexed by compile-time
// constants after unrolling); the tableaus are read from __constant__ memory as direct
// instruction operands.  Self-contained for NVRTC.
#pragma once
#include "sde_common.cuh"
#include "sde_methods_gen.cuh"
#include "sde_series.cuh"

#ifndef SDE_STEP_UNROLL
#define SDE_STEP_UNROLL 1
#endif
// Weight ring of the staged kernels (fixed_body): 0 = register fetch one step ahead (built-in kernels), 1 = cp.async into
// the ring at the end of the previous step (no registers; what the launcher selects for NVRTC systems)
#ifndef SDE_RING_CPASYNC
#define SDE_RING_CPASYNC 0
#endif

namespace sde {

// ------------------------------------------------------------------------------------------
// classic RK4 (src/rk4/gpurk4.jl:73-85).  Quirk Q1: the reference evaluates k1 at ts[i], the END
// of the step being taken; kTimeIsStepEnd makes the caller pass that time.
// ------------------------------------------------------------------------------------------
template <class Sys, class T>
struct RK4Method {
  static constexpr int N = Sys::N;
  static constexpr bool kFSAL = false;
  static constexpr bool kHasExtra = false;
  static constexpr bool kTimeIsStepEnd = true;
  __device__ __forceinline__ void seed(const T*, const T*, T) {}
  __device__ __forceinline__ void begin_step() {}
  template <bool>
  __device__ __forceinline__ void stages(const T* uprev, T* u, const T* p, T t, T dt) {
    const T half = T(0.5);
    const T sixth = T(1) / T(6);
    const T two = T(2);
    T k1[N], k2[N], k3[N], k4[N], tmp[N];
    Sys::rhs(k1, uprev, p, t);
    const T hdt = dt * half;
#pragma unroll
    for (int i = 0; i < N; ++i) tmp[i] = fma(hdt, k1[i], uprev[i]);
    Sys::rhs(k2, tmp, p, fma(half, dt, t));
#pragma unroll
    for (int i = 0; i < N; ++i) tmp[i] = fma(hdt, k2[i], uprev[i]);
    Sys::rhs(k3, tmp, p, fma(half, dt, t));
#pragma unroll
    for (int i = 0; i < N; ++i) tmp[i] = fma(dt, k3[i], uprev[i]);
    Sys::rhs(k4, tmp, p, t + dt);
    const T sdt = dt * sixth;
#pragma unroll
    for (int i = 0; i < N; ++i)
      u[i] = fma(sdt, fma(two, k3[i], fma(two, k2[i], k1[i] + k4[i])), uprev[i]);
  }
  static constexpr int kNB = 1;
  template <bool> __device__ __forceinline__ void dense_prepare(const T*, const T*, T, T) {}
  template <bool> __device__ __forceinline__ void dense_combine(const T*, T, const T*, T*) const {}
  template <bool> __device__ __forceinline__ void dense(T, T, const T*, T*) const {}
};

// ------------------------------------------------------------------------------------------
// forward Euler (src/euler/gpueuler.jl:71-77): t = ts[i] (end of the step, like RK4's quirk Q1),
// u = muladd(dt, k1, uprev).
// ------------------------------------------------------------------------------------------
template <class Sys, class T>
struct EulerMethod {
  static constexpr int N = Sys::N;
  static constexpr bool kFSAL = false;
  static constexpr bool kHasExtra = false;
  static constexpr int kNB = 1;
  __device__ __forceinline__ void seed(const T*, const T*, T) {}
  __device__ __forceinline__ void begin_step() {}
  template <bool>
  __device__ __forceinline__ void stages(const T* uprev, T* u, const T* p, T t, T dt) {
    T k1[N];
    Sys::rhs(k1, uprev, p, t);
#pragma unroll
    for (int i = 0; i < N; ++i) u[i] = fma(dt, k1[i], uprev[i]);
  }
  template <bool> __device__ __forceinline__ void dense_prepare(const T*, const T*, T, T) {}
  template <bool> __device__ __forceinline__ void dense_combine(const T*, T, const T*, T*) const {}
  template <bool> __device__ __forceinline__ void dense(T, T, const T*, T*) const {}
};

// ------------------------------------------------------------------------------------------
// SDE_COMPAT_FAST_STAGES, fixed-step Tsit5: the stage sums with the step size folded into the coefficients,
//     tmp = uprev + sum_j (dt * a_ij) * k_j        instead of the reference's   uprev + dt * (sum_j a_ij * k_j)
// -- h_ij = dt * a_ij is the same for every step and every trajectory of a fixed-step launch, so a stage costs one
// FMA per nonzero coefficient: 21 N instead of 26 N + 1 FP64 instructions per step (Lorenz with its contracted
// right-hand side, SDE_COMPAT_FAST_RHS: 115 -> 99).  NOT the reference's arithmetic: every term is rounded at the magnitude of
// the state; the flag's documented deviation applies (simplediffeq_cuda.h).  Dense output and FSAL handling are the base method's.
// ------------------------------------------------------------------------------------------
template <class Sys, class T>
struct Tsit5FastMethod : Tsit5Method<Sys, T> {
  typedef Tsit5Method<Sys, T> Base;
  static constexpr int N = Sys::N;
  // the launcher's h_ij (KArgs::hcoef: kernel parameters = constant bank, so every h is an FMA operand that costs no
  // register; as 21 per-thread products they took 42 registers -- 94 instead of 60 for the FP64 Lorenz kernel, five
  // CTAs per SM instead of eight, and the shorter loop ran no faster than the reference-exact one)
  const T* h;
  __device__ __forceinline__ void bind(const KArgs<T>& a) { h = a.hcoef; }
  template <bool>
  __device__ __forceinline__ void stages(const T* uprev, T* u, const T* p, T t, T dt) {
    const Tsit5Coef<T>& C = Coefs<T>::tsit5();
    T* k1 = Base::k1; T* k2 = Base::k2; T* k3 = Base::k3; T* k4 = Base::k4; T* k5 = Base::k5; T* k6 = Base::k6; T* k7 = Base::k7;
    T tmp[N];
#pragma unroll
    for (int i = 0; i < N; ++i) tmp[i] = fma(h[0], k1[i], uprev[i]);
    Sys::rhs(k2, tmp, p, fma(C.c1, dt, t));
#pragma unroll
    for (int i = 0; i < N; ++i) tmp[i] = fma(h[2], k2[i], fma(h[1], k1[i], uprev[i]));
    Sys::rhs(k3, tmp, p, fma(C.c2, dt, t));
#pragma unroll
    for (int i = 0; i < N; ++i) tmp[i] = fma(h[5], k3[i], fma(h[4], k2[i], fma(h[3], k1[i], uprev[i])));
    Sys::rhs(k4, tmp, p, fma(C.c3, dt, t));
#pragma unroll
    for (int i = 0; i < N; ++i) tmp[i] = fma(h[9], k4[i], fma(h[8], k3[i], fma(h[7], k2[i], fma(h[6], k1[i], uprev[i]))));
    Sys::rhs(k5, tmp, p, fma(C.c4, dt, t));
#pragma unroll
    for (int i = 0; i < N; ++i)
      tmp[i] = fma(h[14], k5[i], fma(h[13], k4[i], fma(h[12], k3[i], fma(h[11], k2[i], fma(h[10], k1[i], uprev[i])))));
    Sys::rhs(k6, tmp, p, t + dt);
#pragma unroll
    for (int i = 0; i < N; ++i)
      u[i] = fma(h[20], k6[i], fma(h[19], k5[i], fma(h[18], k4[i], fma(h[17], k3[i], fma(h[16], k2[i], fma(h[15], k1[i], uprev[i]))))));
    Sys::rhs(k7, u, p, t + dt);
  }
};
// methods whose coefficients depend on the launch (the step size) bind to the argument block before the step loop
template <class M, class T> __device__ __forceinline__ void method_bind(M&, const KArgs<T>&) {}
template <class Sys, class T> __device__ __forceinline__ void method_bind(Tsit5FastMethod<Sys, T>& m, const KArgs<T>& a) { m.bind(a); }

template <class M> struct MethodTraits { static constexpr bool kTimeIsStepEnd = false; };
template <class Sys, class T> struct MethodTraits<EulerMethod<Sys, T>> { static constexpr bool kTimeIsStepEnd = true; };
template <class Sys, class T> struct MethodTraits<RK4Method<Sys, T>> { static constexpr bool kTimeIsStepEnd = true; };

// ------------------------------------------------------------------------------------------
// output helpers
// ------------------------------------------------------------------------------------------
template <class T, int N>
__device__ __forceinline__ void put_series(const KArgs<T>& a, i64 traj, i64 slot, const T* v) {
  if (a.layout == kLayoutTrajMajor) {
    T* o = a.out_u + (traj * a.n_out + slot) * N;
#pragma unroll
    for (int c = 0; c < N; ++c) o[c] = v[c];
  } else {
#pragma unroll
    for (int c = 0; c < N; ++c) a.out_u[(slot * N + c) * a.ld_out + traj] = v[c];
  }
}

template <class T>
__device__ __forceinline__ void put_series_time(const KArgs<T>& a, i64 traj, i64 slot, T t) {
  if (!a.out_t) return;
  if (a.layout == kLayoutTrajMajor) a.out_t[traj * a.n_out + slot] = t;
  else a.out_t[slot * a.ld_out + traj] = t;
}

template <class T, int N>
__device__ __forceinline__ void put_endpoint(const KArgs<T>& a, i64 traj, const T* v) {
#pragma unroll
  for (int c = 0; c < N; ++c) a.out_u[(i64)c * a.ld_out + traj] = v[c];
}

template <class T, int N, int NP>
__device__ __forceinline__ void load_problem(const KArgs<T>& a, i64 traj, T* u, T* p) {
#pragma unroll
  for (int c = 0; c < N; ++c) u[c] = a.u0[(i64)c * a.ld_in + traj];
#pragma unroll
  for (int c = 0; c < NP; ++c) p[c] = a.p[(i64)c * a.ld_in + traj];
}

// ------------------------------------------------------------------------------------------
// fixed-step body
//   SAVE    kSaveEndpoint | kSaveAt | kSaveEveryStep
//   Q2      reference-exact fixed-step Vern9 dense output (see sde_methods_gen.cuh)
//   STAGED  shared-memory staged trajectory-major series output (see SeriesWriter)
// Follows src/tsit5/gpuatsit5.jl:86-134, src/rk4/gpurk4.jl:66-85, src/verner/gpuvern7.jl:102-228,
// src/verner/gpuvern9.jl:100-339.  The saveat schedule (`while cur_t <= length(ts) && ts[cur_t] <= t`,
// theta and the b_j(theta) polynomials) is the same for every trajectory of a fixed-step solve;
// the launcher evaluates it once on the host with the same IEEE operations (sde_api.cu,
// build_save_plan) and the kernel only does the per-trajectory combination.
// ------------------------------------------------------------------------------------------
template <class Sys, class T, class Method, int SAVE, bool Q2, bool STAGED>
__device__ __forceinline__ void fixed_body(const KArgs<T>& a) {
  constexpr int N = Sys::N, NP = Sys::NP;
  constexpr bool kEnd = MethodTraits<Method>::kTimeIsStepEnd;
  constexpr bool kStaged = STAGED && SAVE != kSaveEndpoint;
  constexpr unsigned FULL = 0xffffffffu;
  i64 traj = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = traj < a.n_traj;
  if (!kStaged) {
    if (!valid) return;
  } else {
    // the staged writer is warp-cooperative: a warp leaves only as a whole, lanes without a trajectory
    // integrate a copy of the last one (and never write)
    if (traj - (i64)(threadIdx.x & 31u) >= a.n_traj) return;
  }
  const i64 src = valid ? traj : a.n_traj - 1;

  T u[N], uprev[N], p[NP > 0 ? NP : 1];
  load_problem<T, N, NP>(a, src, u, p);

  Method m;
  method_bind(m, a);
  T t = a.t0;
  m.seed(u, p, t);
  SeriesWriter<T, N, kStaged> w(a, traj, valid);
  i64 cur = 0;
  if (SAVE == kSaveEveryStep) w.put(u);
  if (SAVE == kSaveAt) {
    if (a.n_save > 0 && a.plan_cnt[0] != 0) {   // us[1] = u0 only when tspan[1] == ts[1] exactly (Q8)
      w.put(u);
      cur = 1;
    }
  }
  // Dense-output weights: plan_b holds NB values per save point with stride NBP (16-byte aligned rows).
  //   direct kernels: vector loads (LDG.128) right where they are used; the SoA kernel is bound by its stores.
  //   staged kernels run few warps per SM next to a large shared-memory carve-out, and a global-memory round trip
  //   per save point was their largest stall (profiles/r2_ncu_trajmajor_staged.txt).  There the warp fetches the
  //   weights of a step ONE STEP AHEAD -- each lane two 16-byte pieces, coalesced, into registers at the top of the
  //   previous step -- parks them in its shared-memory ring at the top of the step, and the save loop reads them
  //   back as warp-wide broadcasts (LDS.128): no global load, no register rotation in the loop.  cp.async straight
  //   into the ring (SDE_RING_CPASYNC) needs no registers but can only start once the previous save loop has released
  //   the ring, i.e. with the stages as its only cover: 6 % slower in the built-in kernels, but 6.5 % FASTER under NVRTC,
  //   whose code spills exactly those registers at the 128-register cap (profiles/r2_tm_variants_line_aligned.txt) --
  //   so the launcher compiles user systems with it.  Save
  //   points beyond the ring's capacity (more than kRingSaves in one step) read their weights from global memory.
  constexpr int kNB = Method::kNB;
  constexpr int kNBP = plan_stride<T>(kNB);
  typedef StageCfg<T, N> SCfg;
  constexpr int kVA = 16 / (int)sizeof(T);
  constexpr int kRingSaves = SCfg::kRingElems / kNBP;                     // save points per ring load
  constexpr int kUnitsPerLane = SCfg::kRingBytes / 16 / 32;               // 16-byte pieces per lane and ring load
  constexpr bool kRing = kStaged && SAVE == kSaveAt;
  static_assert(!kRing || kRingSaves >= 1, "weights of one save point must fit the ring");
  static_assert(kUnitsPerLane == 2, "two 16-byte pieces per lane");
  typedef typename Vec16<T>::type V16;
  const unsigned lane = threadIdx.x & 31u;
  V16 wreg0 = {}, wreg1 = {};     // (two named registers, not an array: NVRTC put a predicated V16[2] into local memory)
  // fetch(first, n): this lane's share of the weights of save points [first, first + n) into registers
  auto fetch = [&](i64 first, int n) {
    const int units = n * (kNBP / kVA);
    const char* gsrc = reinterpret_cast<const char*>(a.plan_b + first * kNBP) + 16u * lane;
    if ((int)lane < units) wreg0 = load16<T>(gsrc);
    if ((int)lane + 32 < units) wreg1 = load16<T>(gsrc + 512);
  };
  auto stash = [&](int n) {
    const int units = n * (kNBP / kVA);
    char* rdst = reinterpret_cast<char*>(w.ring()) + 16u * lane;
    if ((int)lane < units) store16<T>(rdst, wreg0);
    if ((int)lane + 32 < units) store16<T>(rdst + 512, wreg1);
  };
  // the same pieces by cp.async straight into the ring (SDE_RING_CPASYNC)
  auto ring_fetch = [&](i64 first, int n) {
    const int units = n * (kNBP / kVA);
    const char* gsrc = reinterpret_cast<const char*>(a.plan_b + first * kNBP) + 16u * lane;
    char* rdst = reinterpret_cast<char*>(w.ring()) + 16u * lane;
    if ((int)lane < units) async_copy16(rdst, gsrc);
    if ((int)lane + 32 < units) async_copy16(rdst + 512, gsrc + 512);
    async_copy_commit();
  };
  constexpr bool kAsyncRing = SDE_RING_CPASYNC != 0;
  const T dt = a.dt;
  // The number of save points of a step is known BEFORE its stages (a warp-uniform load whose latency hides
  // behind the stages; comparing a prefetched "step of the next save point" after every save point left an L1
  // round trip exposed per save point -- the reason the SoA kernel sat at 0.84-0.90 of the HBM peak in round 1).
  // Staged kernels run the schedule two steps ahead: cnt_next (step s + 1) sizes the weight fetch.
  int cnt_cur = 0, cnt_next = 0;
  if (kRing) {
    if (a.n_steps >= 1) {
      cnt_cur = __shfl_sync(FULL, a.plan_cnt[1], 0);     // the same value, but one ptxas knows to be warp-uniform
      if (kAsyncRing) ring_fetch(cur, cnt_cur < kRingSaves ? cnt_cur : kRingSaves);
      else fetch(cur, cnt_cur < kRingSaves ? cnt_cur : kRingSaves);
    }
    if (a.n_steps >= 2) cnt_next = a.plan_cnt[2];
  }
  // SDE_STEP_UNROLL = 2: two steps per trip, so that ptxas renames registers across the pair instead of copying
  // u -> uprev and (FSAL) k7 -> k1 at the end of every step (12 moves per Lorenz step).  Measured (B200,
  // profiles/r2_saveat_step_unroll_ab.txt): -9 % on the SoA kernel at dt = 0.1 (108 registers: a CTA per SM less),
  // +1.5 % at dt = 0.01 -- off.
  constexpr int kStepUnroll = SDE_STEP_UNROLL;
#pragma unroll(kStepUnroll)
  for (i64 s = 1; s <= a.n_steps; ++s) {
    int cnt = 0;
    if (SAVE == kSaveAt) {
      if (kRing && kAsyncRing) {
        cnt = cnt_cur;                            // (its weights were requested at the end of the previous step)
      } else if (kRing) {
        cnt = cnt_cur;
        __syncwarp();                             // every lane is done with the previous contents of the ring
        stash(cnt < kRingSaves ? cnt : kRingSaves);
        __syncwarp();
        cnt_cur = __shfl_sync(FULL, cnt_next, 0);
        fetch(cur + cnt, cnt_cur < kRingSaves ? cnt_cur : kRingSaves);
        cnt_next = (s + 2 <= a.n_steps) ? a.plan_cnt[s + 2] : 0;
      } else {
        cnt = a.plan_cnt[s];
      }
    }
#pragma unroll
    for (int c = 0; c < N; ++c) uprev[c] = u[c];
    m.begin_step();
    t = kEnd ? a.tgrid[s] : a.tgrid[s - 1];      // range element, never an accumulated sum
    m.template stages<false>(uprev, u, p, t, dt);
    if (!kEnd) t = t + dt;
    if (SAVE == kSaveEveryStep) w.put(u);
    if (SAVE == kSaveAt) {
      if (cnt > 0) {
        m.template dense_prepare<Q2>(uprev, p, t, dt);   // extra stages do not depend on theta: once per step; time base = advanced t (Q3)
        int k = 0;
        if (kRing) {
          if (kAsyncRing) {
            async_copy_wait_all();                // this lane's pieces of the ring have landed ...
            __syncwarp();                         // ... and so have everybody else's
          }
          const int nring = cnt < kRingSaves ? cnt : kRingSaves;
          const T* rb = w.ring();
          for (; k < nring; ++k) {
            T b[kNB];
            load_weights<T, kNB>(rb + k * kNBP, b);
            T o[N];
            m.template dense_combine<Q2>(b, dt, uprev, o);
            w.put(o);
          }
        }
        for (; k < cnt; ++k) {
          T b[kNB];
          load_weights<T, kNB>(a.plan_b + (cur + k) * kNBP, b);
          T o[N];
          m.template dense_combine<Q2>(b, dt, uprev, o);
          w.put(o);
        }
        cur += cnt;
      }
      if (kRing && kAsyncRing) {      // the ring is free: start on the weights of the next step
        cnt_cur = __shfl_sync(FULL, cnt_next, 0);
        __syncwarp();
        ring_fetch(cur, cnt_cur < kRingSaves ? cnt_cur : kRingSaves);
        cnt_next = (s + 2 <= a.n_steps) ? a.plan_cnt[s + 2] : 0;
      }
    }
  }
  if (SAVE == kSaveEndpoint) put_endpoint<T, N>(a, traj, u);
  if (SAVE == kSaveAt) {
    // save points the integration never reached stay `undef` in the reference (quirk Q5): NaN here
    if (cur < a.n_save) {
      T nanv[N];
#pragma unroll
      for (int c = 0; c < N; ++c) nanv[c] = sde_nan(T(0));
      for (; cur < a.n_save; ++cur) w.put(nanv);
    }
  }
  if (SAVE != kSaveEndpoint) w.finish();
}

// ------------------------------------------------------------------------------------------
// adaptive body (PI controller of src/SimpleDiffEq.jl:67-77; loop of src/tsit5/gpuatsit5.jl:250-320,
// src/verner/gpuvern7.jl:353-520, src/verner/gpuvern9.jl:461-763)
//   kV9   AVern9: dtmin / tf-snap thresholds are 1.0f-7 (quirk Q4) and extra-stage times use told
//   kStrict  literal controller arithmetic (pow, divisions, sqrt) instead of the log2-domain one
// ------------------------------------------------------------------------------------------
// `b`, in a form that ptxas can neither relate to the predicate it came from nor evaluate before `x` is known.
// `zero` is 0 at run time (a bit of KArgs::compat that the launcher always clears) but a kernel parameter
// to the compiler.  Why: ptxas hoists the convergence-barrier BREAK of the rejecting lanes to the instruction
// that computes `accept`, and a warp whose lanes disagree then runs the controller code that BOTH outcomes
// share once per group (ncu source counters, shuffled Van der Pol sweep: that block was executed 1.32x per
// attempt; AVern9 at tol 1e-12, 13 % rejections: ~2x).  Branching on this copy moves the divergence behind
// the shared code for 3 instructions: +4 % on the shuffled Van der Pol sweep and on AVern9, -1 % where
// rejections are rare (profiles/r1_adaptive_late_branch_ab.txt).
__device__ __forceinline__ bool late_flag(bool b, double x, int zero) { return b != ((__double2hiint(x) & zero) != 0); }
__device__ __forceinline__ bool late_flag(bool b, float x, int zero) { return b != ((__float_as_int(x) & zero) != 0); }

template <class Sys, class T, class Method, int SAVE, bool kV9, bool kStrict>
__device__ __forceinline__ void adaptive_body(const KArgs<T>& a) {
  constexpr int N = Sys::N, NP = Sys::NP;
  constexpr unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  const double thr = kV9 ? (double)1.0e-7f : 1.0e-14;
  const T beta1 = T(7.0 / 50.0), beta2 = T(2.0 / 25.0), qmax = T(10.0), qmin = T(1.0 / 5.0),
          gamma = T(9.0 / 10.0), qoldinit = T(1.0e-4);
  const T inv_qmax = T(1) / qmax, inv_qmin = T(1) / qmin;
  const T tf = a.tf;

  Method m;
  T u[N], uprev[N], p[NP > 0 ? NP : 1];
  T t = a.t0, dt = a.dt, told = a.t0, dtold = a.dt;
  // controller constants and tables, one shared-memory copy per CTA: k_ctrl (log2-domain controller) or the
  // pow / powf tables of the literal one (k_gpow / k_gpowf; per-lane table indices hit shared memory, not LDG)
  constexpr int kTabCount = kStrict ? (sizeof(T) == 8 ? (int)kGP_count : (int)kGF_count) : (int)kC_count;
  __shared__ __align__(16) double s_ctrl[kTabCount];
  __shared__ __align__(16) T s_bt[Method::kNBT];
  {
    const double* tab_src = kStrict ? (sizeof(T) == 8 ? k_gpow : k_gpowf) : k_ctrl;
    for (int i = threadIdx.x; i < kTabCount; i += blockDim.x) s_ctrl[i] = tab_src[i];
  }
  if (threadIdx.x == 0) Method::load_btilde(s_bt);
  __syncthreads();
  const CtrlTab zlane = s_ctrl;
  // default controller: lqold = beta2 * log2(qold);  literal controller: qoldpow = qold^beta2 (qold itself is
  // used nowhere else, and it only changes on accepted attempts)
  double lqold0 = 0.0;
  T qoldpow0 = T(0);
  if (kStrict) qoldpow0 = strict_exp(beta2, strict_log(qoldinit, zlane), zlane);
  else lqold0 = ctrl_const(zlane, CtrlLog2<T>::kBase + kL_beta2) * ctrl_const(zlane, CtrlLog2<T>::kBase + kL_qoldinit);
  double lqold = lqold0;
  T qoldpow = qoldpow0;
  int cur = 0, nacc = 0, nrej = 0;
  i64 traj = -1;
  bool active = false, drained = false;
  // ---- every accepted step kept (the reference's default, save_everystep = true), trajectory-major rows ----------
  // A lane appending its N values per accepted step to its own row writes 32 different rows per instruction, N * sizeof(T)
  // bytes each (measured, config-1 sweep at 2^20: 5.1x the endpoint-only solve; the NaN fill of the unused capacity went
  // the same way).  Here every lane keeps the open end of its row in a two-line ring of shared memory that is congruent
  // with global memory mod 128 bytes; at the top of every iteration the warp writes out the lines that became complete
  // (four rows per pass, eight lanes per line: LDS.128 -> STG.128), and rows that ended get their last partial line from
  // their owner and their unused capacity (NaN, like the reference's undef) from the whole warp.  Lanes are NOT in
  // lockstep here (work queue), hence ballots instead of the fixed-step writer's warp-uniform counters.
  // Needs N * sizeof(T) <= 64: a row grows by at most two slots between two services (u0 and the first accepted step of a
  // refilled lane), which must fit next to a pending partial line.  Times go straight to out_t (8 bytes per step).
  constexpr bool kRowStage = (SAVE == kSaveEveryStep) && (N * (int)sizeof(T) <= kRowStageMaxSlotBytes);
  const bool rowstage = kRowStage && a.layout == kLayoutTrajMajor && blockDim.x >= 32;
  constexpr int kLineE = 128 / (int)sizeof(T);
  constexpr int kRingE = 2 * kLineE;
  T* rs_ring = reinterpret_cast<T*>(sde_dyn_smem + (size_t)threadIdx.x * kRowStageStrideB);
  int rs_wpos = 0;          // next element of the row stream (element 0 = start of the line that holds the row's start)
  int rs_lines = 0;         // lines of the row stream already in global memory
  i64 rs_end = 0;           // end of the row in the same coordinates
  char* rs_gline = nullptr; // global address of row-stream element 0
  bool rs_open = false, rs_fin = false;
  auto row_append = [&](const T* v) {
#pragma unroll
    for (int c = 0; c < N; ++c) rs_ring[(rs_wpos + c) & (kRingE - 1)] = v[c];
    rs_wpos += N;
  };
  auto row_phi = [&]() {    // offset of the row's start in its first line, in elements (recomputed: once or twice per row)
    return (int)(((u64)a.out_u + (u64)(traj * a.n_out * N) * sizeof(T)) & (u64)127) / (int)sizeof(T);
  };
  auto row_service = [&]() {
    // (1) lines that became complete since the last service (at most one per lane).  Every ready lane publishes
    // where its line lies (shared-memory offset) and where it goes (global address) under its rank among the ready
    // lanes; then eight lanes move one line, four lines per pass.
    const bool ready = rs_open && (rs_wpos / kLineE > rs_lines);
    const unsigned mask = __ballot_sync(FULL, ready);
    if (mask) {
      unsigned char* wbase = sde_dyn_smem + (size_t)(threadIdx.x & ~31u) * kRowStageStrideB;
      RowLineMeta* meta = reinterpret_cast<RowLineMeta*>(sde_dyn_smem + (size_t)blockDim.x * kRowStageStrideB) + (threadIdx.x & ~31u);
      if (ready) {
        RowLineMeta mt;
        // (line 0 is shared with the previous row: its owner writes its part below, the warp skips it)
        mt.dst = rs_lines > 0 ? (u64)(rs_gline + (size_t)rs_lines * 128) : 0;
        mt.src = (unsigned)(lane * kRowStageStrideB + (rs_lines & 1) * 128);
        mt.pad = 0;
        meta[__popc(mask & ((1u << lane) - 1u))] = mt;
      }
      __syncwarp();
      const int nready = __popc(mask);
      const unsigned piece = (lane & 7u) * 16u;
      for (int k = (int)(lane >> 3); k < nready; k += 4) {
        const RowLineMeta mt = meta[k];
        if (mt.dst) copy16<T>(reinterpret_cast<char*>(mt.dst) + piece, wbase + mt.src + piece);
      }
      __syncwarp();
      if (ready) {
        if (rs_lines == 0) {
          T* g = reinterpret_cast<T*>(rs_gline);
          for (int e = row_phi(); e < kLineE; ++e) g[e] = rs_ring[e];
        }
        ++rs_lines;
      }
    }
    // (2) rows that ended: the pending partial line by its owner, the unused capacity by the warp
    unsigned fmask = __ballot_sync(FULL, rs_fin);
    if (fmask) {
      if (rs_fin) {
        T* g = reinterpret_cast<T*>(rs_gline);
        for (int e = (rs_lines == 0) ? row_phi() : rs_lines * kLineE; e < rs_wpos; ++e) g[e] = rs_ring[e & (kRingE - 1)];
      }
      const T nan = sde_nan(T(0));
      while (fmask) {
        const int r = __ffs(fmask) - 1;
        fmask &= fmask - 1u;
        T* g = reinterpret_cast<T*>(__shfl_sync(FULL, (u64)rs_gline, r));
        const i64 from = (i64)__shfl_sync(FULL, rs_wpos, r), to = __shfl_sync(FULL, rs_end, r);
        for (i64 e = from + lane; e < to; e += 32) g[e] = nan;
        const i64 tr = __shfl_sync(FULL, traj, r);
        const int na = __shfl_sync(FULL, nacc, r);
        if (a.out_t) {
          T* gt = a.out_t + tr * a.n_out;
          for (i64 sl = (i64)na + 1 + lane; sl < a.n_out; sl += 32) gt[sl] = nan;
        }
      }
      if (rs_fin) { rs_fin = false; rs_open = false; }
    }
  };
  // attempts are counted as nacc + nrej (int32, like the naccept / nreject outputs); 0 = the
  // reference's "no maxiters" = the int32 range
  const int attempt_limit = (a.max_attempts > 0 && a.max_attempts < (i64)0x7fffffff) ? (int)a.max_attempts : 0x7fffffff;

  for (;;) {
    if (rowstage) row_service();
    // ---- work queue: idle lanes fetch the next trajectory (one atomic per warp per refill).
    // Fast path while every lane is busy: one vote.
    if (__any_sync(FULL, !active)) {
      const unsigned want = __ballot_sync(FULL, !active && !drained);
      if (want) {
        const int leader = __ffs(want) - 1;
        u64 base = 0;
        if ((int)lane == leader) base = atomicAdd(a.queue, (u64)__popc(want));
        base = __shfl_sync(FULL, base, leader);
        if (!active && !drained) {
          traj = (i64)base + __popc(want & ((1u << lane) - 1u));
          if (traj < a.n_traj) {
            load_problem<T, N, NP>(a, traj, u, p);
            t = a.t0; dt = a.dt; told = a.t0; dtold = a.dt;
            lqold = lqold0;
            qoldpow = qoldpow0;
            cur = 0; nacc = 0; nrej = 0;
            m.seed(u, p, t);
  #pragma unroll
            for (int c = 0; c < N; ++c) uprev[c] = u[c];
            m.begin_step();
            if (SAVE == kSaveAt) {
              if (a.n_save > 0 && a.t0 == a.saveat[0]) {
                put_series<T, N>(a, traj, 0, u);
                cur = 1;
              }
            }
            if (SAVE == kSaveEveryStep) {   // us = [u0], ts = [t0]   (gpuatsit5.jl:220-224)
              if (rowstage) {
                const u64 mine = (u64)a.out_u + (u64)(traj * a.n_out * N) * sizeof(T);
                rs_gline = reinterpret_cast<char*>(mine & ~(u64)127);
                rs_wpos = (int)(mine & (u64)127) / (int)sizeof(T);
                rs_end = (i64)rs_wpos + a.n_out * N;
                rs_lines = 0;
                rs_open = true;
                row_append(u);
              } else {
                put_series<T, N>(a, traj, 0, u);
              }
              put_series_time<T>(a, traj, 0, t);
            }
            active = true;
            if (!(t < tf)) {   // `while t < tspan[2]` never entered
              if (rowstage) rs_fin = true;      // (the row still gets its slot 0 and its NaN capacity)
              if (SAVE == kSaveEndpoint) put_endpoint<T, N>(a, traj, u);
              if (SAVE != kSaveEveryStep && a.out_t) a.out_t[traj] = t;
              if (a.naccept) a.naccept[traj] = 0;
              if (a.nreject) a.nreject[traj] = 0;
              if (a.retcode) a.retcode[traj] = kRetDefault;
              active = false;
            }
          } else {
            drained = true;
          }
        }
      }
      if (__all_sync(FULL, !active)) {
        if (__all_sync(FULL, drained)) {        // warp-vote exit: nothing left anywhere
          if (rowstage) row_service();          // (a row that ended inside this refill)
          break;
        }
        continue;
      }
    }

    if (active) {
      bool fin = false;          // trajectory finished (or failed) in this attempt
      int ret = kRetDefault;
      if ((double)dt < thr) {
        fin = true; ret = kRetDtMin;                   // error("dt<dtmin")
      } else if (nacc + nrej >= attempt_limit) {
        fin = true; ret = kRetMaxIters;
      } else {
        m.template stages<true>(uprev, u, p, t, dt);
        T e[N];
        m.error(dt, e, s_bt);
        bool accept;
        T EEst = T(0);                 // literal controller only: the estimate and the log half of its powers
        decltype(strict_log(T(1), zlane)) Elog{};
        if (kStrict) {
          // the reference's arithmetic, operation for operation (gpuatsit5.jl:276-292):
          // tmp ./ (abstol .+ max.(abs.(uprev), abs.(u)) * reltol) ; ODE_DEFAULT_NORM
          // (divisions: strict_div() = the fast path of CUDA's own IEEE division without its branch, sde_common.cuh;
          //  a group whose range tests all pass is done, otherwise the group is recomputed with plain divisions)
          T den[N], sc[N];
#pragma unroll
          for (int c = 0; c < N; ++c) den[c] = a.abstol + max_abs_nan2(uprev[c], u[c]) * a.reltol;
          bool ok = true;
#pragma unroll
          for (int c = 0; c < N; ++c) sc[c] = strict_div(e[c], den[c], ok);
          if (!ok) {
#pragma unroll
            for (int c = 0; c < N; ++c) sc[c] = e[c] / den[c];
          }
          if (N == 1) {
            EEst = sde_abs(sc[0]);
          } else {
            T ssum = T(0);
#pragma unroll
            for (int c = 0; c < N; ++c) ssum = (c == 0) ? sc[c] * sc[c] : ssum + sc[c] * sc[c];
            ok = true;
            T mean = strict_div(ssum, T(N), ok);
            if (!ok) mean = ssum / T(N);
            EEst = sde_sqrt(mean);
          }
          accept = !(EEst > T(1));
          // q11 = EEst^beta1: log half once (kept for qold^beta2 below), exp half per exponent; anything but a
          // positive normal estimate (0, subnormal, inf, NaN) takes the out-of-line complete function
          Elog = strict_log(EEst, zlane);      // (table indices are masked: harmless for any bit pattern)
          T q11 = strict_exp(beta1, Elog, zlane);
          if (!strict_is_main(EEst)) q11 = sde_pow_cold(EEst, beta1);
          // accept: q = EEst == 0 ? inv(qmax) : q11 / qold^beta2;  q = max(inv(qmax), min(inv(qmin), q / gamma));  dt /= q
          // reject: dt /= min(inv(qmin), q11 / gamma)     (EEst > 1, so q11 / gamma > 1: the lower clamp is a no-op and
          //         the NaN rule of Base.min is never exercised)
          // both as straight-line code: the lanes of a warp part only at the late branch below
          ok = true;
          T q = accept ? strict_div(q11, qoldpow, ok) : q11;
          if (EEst == T(0)) q = inv_qmax;
          q = max_fast(inv_qmax, min_fast(inv_qmin, strict_div(q, gamma, ok)));
          T dtnew = strict_div(dt, q, ok);
          if (!ok) {
            q = accept ? q11 / qoldpow : q11;
            if (EEst == T(0)) q = inv_qmax;
            q = max_fast(inv_qmax, min_fast(inv_qmin, q / gamma));
            dtnew = dt / q;
          }
          if (accept) dtold = dt;
          dt = dtnew;
        } else {
          // same formulas in the log2 domain (see sde_common.cuh); lqold = beta2 * log2(qold)
          constexpr int LB = CtrlLog2<T>::kBase;
          const double one = ctrl_const(zlane, kC_one);
          double lE;        // log2(EEst)
          bool ezero;       // the reference's `EEst == 0 ? inv(qmax)` branch: q / gamma = 1/9, NOT the clamp's 1/10
          if (sizeof(T) == 8) {
            // EEst^2 = sum((e_i / sc_i)^2) / N with Newton reciprocals: no division, no sqrt
            double ss = 0.0;
#pragma unroll
            for (int c = 0; c < N; ++c) {
              const double sc = (double)(a.abstol + max_abs_nan2(uprev[c], u[c]) * a.reltol);
              const double x = (double)e[c] * sde_rcp_fast(sc, one);
              ss = (c == 0) ? x * x : ss + x * x;
            }
            ss = ss * (1.0 / (double)N);
            accept = !(ss > one);
            ezero = (ss == 0.0);
            lE = (ss != ss) ? ss : ctrl_const(zlane, kC_half) * sde_log2_fast(ss, zlane);
          } else {
            // FP32 state: EEst exactly as the reference (IEEE float div / sqrt), controller in FP64
            if (N == 1) {
              EEst = sde_abs(e[0] / (a.abstol + max_abs_nan2(uprev[0], u[0]) * a.reltol));
            } else {
              T ssum = T(0);
#pragma unroll
              for (int c = 0; c < N; ++c) {
                const T sc = e[c] / (a.abstol + max_abs_nan2(uprev[c], u[c]) * a.reltol);
                ssum = (c == 0) ? sc * sc : ssum + sc * sc;
              }
              EEst = sde_sqrt(ssum / T(N));
            }
            accept = !(EEst > T(1));
            ezero = (EEst == T(0));
            lE = (EEst != EEst) ? (double)EEst : sde_log2_fast((double)EEst, zlane);
          }
          // reject:  dt /= min(inv(qmin), q11/gamma)                     -> exponent -min(l_invqmin, l11 - l_gamma)
          // accept:  q = max(inv(qmax), min(inv(qmin), q11/qold^beta2/gamma)); dt /= q   -> exponent -lq
          // (EEst > 1 on the reject path, so its min needs no NaN rule; the accept clamp uses the
          //  FastMath forms, which map a NaN estimate to inv(qmax) like the reference)
          const double l11 = ctrl_const(zlane, LB + kL_beta1) * lE;
          const double lg = ctrl_const(zlane, LB + kL_gamma);
          const double lmin = ctrl_const(zlane, LB + kL_inv_qmin);
          const double lmax = ctrl_const(zlane, LB + kL_inv_qmax);
          // (on the reject path l11 > 0 > lmax + lg, so the lower clamp is a no-op there and is applied
          //  unconditionally)
          double lx = accept ? l11 - lqold : l11;
          if (ezero) lx = lmax;            // (log2(0) is about -1023 here; the clamp alone would give 1/qmax instead of 1/qmax/gamma)
          lx = lx - lg;
          lx = max_fast(lmax, min_fast(lmin, lx));
          if (accept) {
            const double lq0 = ctrl_const(zlane, LB + kL_qoldinit);
            lqold = ctrl_const(zlane, LB + kL_beta2) * ((lE > lq0) ? lE : lq0);   // qold = max(EEst, qoldinit)
            dtold = dt;
          }
          dt = (T)((double)dt * sde_exp2_fast(-lx, zlane));
        }
        if (!late_flag(accept, dt, a.compat & kCompatRuntimeZero)) {
          ++nrej;
        } else {
          const T rem = tf - t - dtold;
          dt = min_abs_nan1(dt, rem);        // min(abs(dt), abs(tf - t - dtold))
          told = t;
          if ((double)rem < thr) t = tf;
          else t = t + dtold;
          ++nacc;
          if (kStrict) {
            // qold = max(EEst, qoldinit), carried as qold^beta2.  qoldinit < EEst <= 1 is a positive normal number;
            // Base.max propagates a NaN estimate
            if (EEst > qoldinit) qoldpow = strict_exp(beta2, Elog, zlane);
            else qoldpow = (EEst != EEst) ? EEst : qoldpow0;
          }
          if (SAVE == kSaveEveryStep) {   // push!(us, u); push!(ts, t)   (gpuatsit5.jl:301-303)
            if ((i64)nacc < a.n_out) {
              if (rowstage) row_append(u);
              else put_series<T, N>(a, traj, nacc, u);
              put_series_time<T>(a, traj, nacc, t);
            }
          }
          if (SAVE == kSaveAt) {
            bool prepared = false;
            while (cur < a.n_save && a.saveat[cur] <= t) {
              const T savet = a.saveat[cur];
              const T th = (savet - told) / dtold;
              if (!prepared) {
                m.template dense_prepare<false>(uprev, p, kV9 ? told : t, dtold);
                prepared = true;
              }
              T o[N];
              m.template dense<false>(th, dtold, uprev, o);
              put_series<T, N>(a, traj, cur, o);
              ++cur;
            }
          }
          fin = !(t < tf);
          // the next step starts from the accepted state (after the dense output, which reads the
          // old uprev / k1)
#pragma unroll
          for (int c = 0; c < N; ++c) uprev[c] = u[c];
          m.begin_step();
        }
      }
      if (fin) {   // publish and free the lane
        if (SAVE == kSaveEndpoint) put_endpoint<T, N>(a, traj, u);
        if (SAVE == kSaveAt && cur < a.n_save) {   // never reached (failure, or saveat beyond tf)
          T nanv[N];
#pragma unroll
          for (int c = 0; c < N; ++c) nanv[c] = sde_nan(T(0));
          for (; cur < a.n_save; ++cur) put_series<T, N>(a, traj, cur, nanv);
        }
        if (SAVE == kSaveEveryStep) {
          if (ret == kRetDefault && (i64)nacc >= a.n_out) ret = kRetOutputFull;
          if (rowstage) {
            rs_fin = true;              // last partial line + unused capacity: row_service() at the top of the next iteration
          } else {
            T nanv[N];
#pragma unroll
            for (int c = 0; c < N; ++c) nanv[c] = sde_nan(T(0));
            for (i64 s = (i64)nacc + 1; s < a.n_out; ++s) {   // unused capacity
              put_series<T, N>(a, traj, s, nanv);
              put_series_time<T>(a, traj, s, nanv[0]);
            }
          }
        } else if (a.out_t) {
          a.out_t[traj] = t;
        }
        if (a.naccept) a.naccept[traj] = nacc;
        if (a.nreject) a.nreject[traj] = nrej;
        if (a.retcode) a.retcode[traj] = ret;
        active = false;
      }
    }
  }
}

}  // namespace sde
