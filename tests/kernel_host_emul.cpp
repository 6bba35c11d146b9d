// Host emulation of the DEVICE kernel bodies (csrc/device/sde_kernels.cuh: fixed_body, adaptive_body), so that
// the CPU test suite can run the very source the GPU runs -- stage code, save handling, work queue, step
// controller, late accept branch -- against the oracle without a GPU.
//
// *** TEST INFRASTRUCTURE ONLY ***  Built and loaded by tests/test_kernel_host_emul.py; nothing in the product
// links it and it is no CPU fallback: the library itself still fails with SDE_ERR_CUDA without a device.
//
// Emulation model, two modes (emul_set_lanes):
//   1 lane : one thread per "warp" and per block; the warp intrinsics degenerate (ballot = bit 0, shuffle =
//            identity, votes = the lane's own predicate).  Fast; used for most tests.
//   32 lanes: one block = one warp of 32 host threads.  Every warp intrinsic the kernels use is called by all
//            lanes of the warp at the same program point (the kernels keep those calls out of divergent code), so
//            each one is a rendezvous: deposit the operand, barrier, read the others', barrier.  __shared__ arrays
//            are shared by the block, atomicAdd is a real atomic.  This runs the warp-aggregated work queue
//            (ballot / popc prefix / leader atomic / shuffle), the warp-vote exit and the shared-memory STAGED
//            trajectory-major writer (whole-line flushes and the per-warp weight ring between __syncwarp()s) as 32 cooperating lanes.
// Blocks run one after the other.  The only arithmetic that differs
// from the device is the seed of sde_rcp_fast: MUFU.RCP64H there, the IEEE quotient 1.0 / x here (the float seed
// that SDE_HOST_EMULATION selects in sde_common.cuh overflows for |x| > 3.4e38, which blown-up trajectories reach;
// the device instruction covers the whole double range).  Both are refined by the same two Newton steps.
// Fixed-step kernels do not use it and must match the oracle bit for bit, adaptive kernels are held to the same
// bar as on the GPU (identical step counts, states within tolerance).
#include <cmath>
#include <cstdint>
#include <cstring>

// sde_rcp_fast's one inline-PTX statement, `asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));`, becomes the
// assignment below (r and x are the names in scope there; no other asm statement exists in the included headers)
#define asm(...) (r = 1.0 / x)
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __constant__
#define __shared__ static      /* shared by the block's threads; blocks run one after the other */
// The headers' one non-local shared declaration, `extern __shared__ __align__(16) unsigned char sde_dyn_smem[];`
// (dynamic shared memory of the staged writer), cannot become `extern static`: the test fixture compiles against
// copies of the device headers in which exactly that line reads EMUL_DYN_SMEM (nothing else is touched).
#define EMUL_DYN_SMEM static __attribute__((aligned(16))) unsigned char sde_dyn_smem[65536];
#define __restrict__
#define __align__(n) __attribute__((aligned(n)))
struct double2 { double x, y; };
struct EmulDim { unsigned x = 0, y = 0, z = 0; };
static thread_local EmulDim threadIdx, blockIdx;
static thread_local EmulDim blockDim, gridDim;

static inline int __double2hiint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo) {
  int64_t b = (int64_t)(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo); double x; std::memcpy(&x, &b, 8); return x;
}
static inline double __longlong_as_double(long long v) { double x; std::memcpy(&x, &v, 8); return x; }
static inline long long __double_as_longlong(double x) { long long v; std::memcpy(&v, &x, 8); return v; }
static inline float __int_as_float(int v) { float x; std::memcpy(&x, &v, 4); return x; }
static inline int __float_as_int(float v) { int x; std::memcpy(&x, &v, 4); return x; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
// ---- warp of g_lanes host threads (1 or 32) --------------------------------------------------------------------
#include <pthread.h>
static int g_lanes = 1;
static pthread_barrier_t g_bar;
static unsigned long long g_slot[32];
static inline void warp_barrier() { if (g_lanes > 1) pthread_barrier_wait(&g_bar); }
static inline unsigned __ballot_sync(unsigned, int p) {
  if (g_lanes == 1) return p ? 1u : 0u;
  g_slot[threadIdx.x & 31u] = p ? 1ull : 0ull;
  warp_barrier();
  unsigned b = 0;
  for (int l = 0; l < g_lanes; ++l) b |= (unsigned)g_slot[l] << l;
  warp_barrier();
  return b;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0u; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, p) == (g_lanes == 1 ? 1u : 0xffffffffu); }
template <class V> static inline V __shfl_sync(unsigned, V v, int src, int = 32) {
  if (g_lanes == 1) return v;
  static_assert(sizeof(V) <= 8, "shuffle operand");
  unsigned long long raw = 0;
  std::memcpy(&raw, &v, sizeof(V));
  g_slot[threadIdx.x & 31u] = raw;
  warp_barrier();
  raw = g_slot[src & 31];
  warp_barrier();
  V out;
  std::memcpy(&out, &raw, sizeof(V));
  return out;
}
static inline void __syncthreads() { warp_barrier(); }      // one warp per block
static inline void __syncwarp(unsigned = 0xffffffffu) { warp_barrier(); }
template <class V> static inline V atomicAdd(V* p, V v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
using std::fma;
using std::fabs;
using std::sqrt;

#include <thread>
#include <vector>

static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * (uint64_t)b) >> 32); }
// FP32 Box-Muller of the EM generator: CUDA's sincospif has no libm twin; the FP32 stream is only held to 3e-6
static inline void sincospif(float x, float* s, float* c) {
  const double a = 3.14159265358979323846 * (double)x;
  *s = (float)std::sin(a); *c = (float)std::cos(a);
}

#include "sde_kernels.cuh"        // copies of csrc/device/*.cuh made by the test fixture (see EMUL_DYN_SMEM below)
#include "sde_systems.cuh"
#include "sde_em.cuh"
#include "sde_interp_host_gen.h"   // csrc/: the launcher's dense-output polynomial tables


namespace {

struct Call {
  int alg, save, compat, layout;
  long long n_traj, n_steps, n_save, n_out, max_attempts;
  double t0, tf, dt, abstol, reltol;
  const void *u0, *p, *tgrid, *saveat;
  void *out_u, *out_t;
  int *naccept, *nreject, *retcode;
};

// run `body` as n_blocks blocks of g_lanes threads, one block after the other
template <class Body>
void launch_blocks(long long n_blocks, Body body) {
  for (long long b = 0; b < n_blocks; ++b) {
    if (g_lanes == 1) {
      blockDim.x = 1; gridDim.x = (unsigned)n_blocks; blockIdx.x = (unsigned)b; threadIdx.x = 0;
      body();
      continue;
    }
    std::vector<std::thread> lanes;
    for (int l = 0; l < g_lanes; ++l)
      lanes.emplace_back([=]() {
        blockDim.x = (unsigned)g_lanes; gridDim.x = (unsigned)n_blocks; blockIdx.x = (unsigned)b; threadIdx.x = (unsigned)l;
        body();
      });
    for (auto& t : lanes) t.join();
  }
}

template <class T>
sde::KArgs<T> make_args(const Call& c, sde::u64* queue) {
  sde::KArgs<T> a;
  std::memset(&a, 0, sizeof a);
  a.u0 = (const T*)c.u0; a.p = (const T*)c.p;
  a.n_traj = c.n_traj; a.ld_in = c.n_traj;
  a.t0 = (T)c.t0; a.tf = (T)c.tf; a.dt = (T)c.dt; a.abstol = (T)c.abstol; a.reltol = (T)c.reltol;
  a.n_steps = c.n_steps; a.tgrid = (const T*)c.tgrid; a.saveat = (const T*)c.saveat; a.n_save = (int)c.n_save;
  a.compat = c.compat & 3;           // like the launcher: bit 30 reaches the kernel as 0
  a.layout = c.layout; a.max_attempts = c.max_attempts;
  a.out_u = (T*)c.out_u; a.ld_out = c.n_traj; a.n_out = c.n_out; a.out_t = (T*)c.out_t;
  a.naccept = c.naccept; a.nreject = c.nreject; a.retcode = c.retcode;
  a.queue = queue;
  // like the launcher (sde_api.cu) for SDE_COMPAT_FAST_STAGES: h_ij = dt * a_ij
  for (int k = 0; k < sde_host::kTsit5NStageCoef; ++k) a.hcoef[k] = (T)((T)c.dt * (T)sde_host::kTsit5StageCoef[k]);
  return a;
}

// The fixed-step save schedule is built on the host by the launcher (sde_api.cu: build_save_plan) and handed to the
// kernel as plan_step / plan_b.  Restated here with the same operations (t = tgrid[s-1] + dt; theta = (savet -
// (t - dt)) / dt; b_j(theta) by fma-Horner over the SAME tables, sde_interp_host_gen.h) so that the kernel's side --
// the save loop, dense_prepare (extra stages, quirks Q2 / Q3) and dense_combine -- can be run here.
template <class T>
void build_plan(int alg, const T* tgrid, long long n_steps, T t0, T dt, const T* saveat, long long n_save,
                std::vector<int>* step, std::vector<T>* b) {
  const double* poly; const int* len; int nb, deg;
  if (alg == sde::kTsit5) { poly = &sde_host::kTsit5Poly[0][0]; len = sde_host::kTsit5Len; nb = sde_host::kTsit5NB; deg = sde_host::kTsit5Deg; }
  else if (alg == sde::kVern7) { poly = &sde_host::kVern7Poly[0][0]; len = sde_host::kVern7Len; nb = sde_host::kVern7NB; deg = sde_host::kVern7Deg; }
  else { poly = &sde_host::kVern9Poly[0][0]; len = sde_host::kVern9Len; nb = sde_host::kVern9NB; deg = sde_host::kVern9Deg; }
  const int va = (int)(16 / sizeof(T)), nbp = (nb + va - 1) / va * va;   // row stride of the weights (sde::plan_stride)
  step->assign((size_t)n_save, (int)(n_steps + 1));      // "never reached"
  b->assign((size_t)n_save * nbp + va, (T)0);
  long long cur = 0;
  if (n_save > 0 && t0 == saveat[0]) { (*step)[0] = 0; cur = 1; }
  for (long long s = 1; s <= n_steps && cur < n_save; ++s) {
    volatile T tv = tgrid[s - 1];
    tv = tv + dt;
    const T t = tv;
    while (cur < n_save && saveat[cur] <= t) {
      volatile T tm = t - dt;
      const T th = (saveat[cur] - tm) / dt;
      for (int j = 0; j < nb; ++j) {
        const double* c = poly + (size_t)j * deg;
        T acc = (T)c[len[j] - 1];
        for (int d = len[j] - 2; d >= 0; --d) acc = std::fma(th, acc, (T)c[d]);
        (*b)[(size_t)cur * nbp + j] = acc;
      }
      (*step)[(size_t)cur] = (int)s;
      ++cur;
    }
  }
}

// fixed step: the grid is one thread per trajectory
template <class Sys, class T, class M, int SAVE, bool Q2 = false>
void run_fixed(const Call& c) {
  sde::KArgs<T> a = make_args<T>(c, nullptr);
  std::vector<int> plan_step;
  std::vector<T> plan_b;
  if (SAVE == sde::kSaveAt) {
    build_plan<T>(c.alg, (const T*)c.tgrid, c.n_steps, (T)c.t0, (T)c.dt, (const T*)c.saveat, c.n_save, &plan_step, &plan_b);
    std::vector<int> cnt((size_t)c.n_steps + 1, 0);        // save points per step, like the launcher (sde_api.cu)
    for (int st : plan_step) if (st >= 0 && (long long)st <= c.n_steps) ++cnt[(size_t)st];
    plan_step.swap(cnt);
    a.plan_cnt = plan_step.data();
    a.plan_b = plan_b.data();
  }
  const long long n_blocks = (c.n_traj + g_lanes - 1) / g_lanes;
  if constexpr (SAVE != sde::kSaveEndpoint) {
    if (c.compat & 16) {      // emulation-only flag: the STAGED trajectory-major writer (needs the 32-lane mode)
      launch_blocks(n_blocks, [&]() { sde::fixed_body<Sys, T, M, SAVE, Q2, true>(a); });
      return;
    }
  }
  launch_blocks(n_blocks, [&]() { sde::fixed_body<Sys, T, M, SAVE, Q2, false>(a); });
}

// adaptive: one persistent thread drains the whole work queue
template <class Sys, class T, class M, int SAVE, bool V9>
void run_adaptive(const Call& c) {
  sde::u64 queue[2] = {0, 0};
  sde::KArgs<T> a = make_args<T>(c, queue);
  // persistent blocks: the first one drains the queue, the second finds it empty and leaves through the vote exit
  if (c.compat & 2) launch_blocks(2, [&]() { sde::adaptive_body<Sys, T, M, SAVE, V9, true>(a); });
  else launch_blocks(2, [&]() { sde::adaptive_body<Sys, T, M, SAVE, V9, false>(a); });
}

template <class Sys, class T>
int dispatch_alg(const Call& c) {
  using namespace sde;
  using TS = Tsit5Method<Sys, T>; using RK = RK4Method<Sys, T>; using EU = EulerMethod<Sys, T>;
  using V7 = Vern7Method<Sys, T>; using V9 = Vern9Method<Sys, T>;
#define FIX(M) do { if (c.save == kSaveEndpoint) run_fixed<Sys, T, M, kSaveEndpoint>(c); \
                    else if (c.save == kSaveEveryStep) run_fixed<Sys, T, M, kSaveEveryStep>(c); else return -4; return 0; } while (0)
#define FIXS(M) do { if (c.save == kSaveEndpoint) run_fixed<Sys, T, M, kSaveEndpoint>(c); \
                     else if (c.save == kSaveEveryStep) run_fixed<Sys, T, M, kSaveEveryStep>(c); \
                     else run_fixed<Sys, T, M, kSaveAt>(c); return 0; } while (0)
#define ADA(M, V) do { if (c.save == kSaveEndpoint) run_adaptive<Sys, T, M, kSaveEndpoint, V>(c); \
                       else if (c.save == kSaveAt) run_adaptive<Sys, T, M, kSaveAt, V>(c); \
                       else run_adaptive<Sys, T, M, kSaveEveryStep, V>(c); return 0; } while (0)
  switch (c.alg) {
    case kTsit5: FIXS(TS);
    case kRK4: FIX(RK);
    case kEuler: FIX(EU);
    case kVern7: FIXS(V7);
    case kVern9:      // saveat: the reference's dense output (quirk Q2) unless SDE_COMPAT_FIX_VERN9_INTERP (bit 0)
      if (c.save == kSaveAt) { if (c.compat & 1) run_fixed<Sys, T, V9, kSaveAt, false>(c); else run_fixed<Sys, T, V9, kSaveAt, true>(c); return 0; }
      FIX(V9);
    case kATsit5: ADA(TS, false);
    case kAVern7: ADA(V7, false);
    case kAVern9: ADA(V9, true);
  }
#undef FIX
#undef FIXS
#undef ADA
  return -1;
}

template <class T>
int dispatch_sys(int sys, const Call& c) {
  switch (sys) {
    case 0: return dispatch_alg<sde::Lorenz, T>(c);
    case 1: return dispatch_alg<sde::VanDerPol, T>(c);
    case 2: return dispatch_alg<sde::Robertson, T>(c);
    case 4: return dispatch_alg<sde::LinearDecay, T>(c);
    case 5: return dispatch_alg<sde::ScalarGrowth, T>(c);
    case 6: return dispatch_alg<sde::NonAutonomous, T>(c);
  }
  return -1;   // (3 = nbody: left out to keep this translation unit small; its kernels are covered on the GPU)
}

}  // namespace

// sys: index into the library's registry order (lorenz, vanderpol, robertson, nbody, lineardecay, scalargrowth,
// nonautonomous); the other arguments are the fields of sde::KArgs / sde_options_t with the same meaning.
extern "C" int emul_solve(int sys, int alg, int dtype, int save, int layout, int compat, long long n_traj,
                          const void* u0, const void* p, double t0, double tf, double dt, double abstol, double reltol,
                          long long n_steps, const void* tgrid, const void* saveat, long long n_save, long long n_out,
                          long long max_attempts, void* out_u, void* out_t, int* naccept, int* nreject, int* retcode) {
  Call c{alg, save, compat, layout, n_traj, n_steps, n_save, n_out, max_attempts, t0, tf, dt, abstol, reltol,
         u0, p, tgrid, saveat, out_u, out_t, naccept, nreject, retcode};
  return dtype == 0 ? dispatch_sys<double>(sys, c) : dispatch_sys<float>(sys, c);
}

// ---- SDE_COMPAT_FAST_RHS, fixed-step Tsit5, endpoint only: the contracted right-hand-side twins with the step size folded
// into the stage coefficients (sde_kernels.cuh: Tsit5FastMethod) -- the kernels sde_builtin.cuh instantiates for them.
// sys: 0 = lorenz twin, 1 = vanderpol twin
extern "C" int emul_fast_tsit5(int sys, int dtype, long long n_traj, const void* u0, const void* p, double t0, double dt,
                               long long n_steps, const void* tgrid, void* out_u) {
  Call c{sde::kTsit5, sde::kSaveEndpoint, 0, 0, n_traj, n_steps, 0, 1, 0, t0, 0.0, dt, 0.0, 0.0,
         u0, p, tgrid, nullptr, out_u, nullptr, nullptr, nullptr, nullptr};
  if (sys == 0 && dtype == 0) run_fixed<sde::LorenzFma, double, sde::Tsit5FastMethod<sde::LorenzFma, double>, sde::kSaveEndpoint>(c);
  else if (sys == 0) run_fixed<sde::LorenzFma, float, sde::Tsit5FastMethod<sde::LorenzFma, float>, sde::kSaveEndpoint>(c);
  else if (sys == 1 && dtype == 0) run_fixed<sde::VanDerPolFma, double, sde::Tsit5FastMethod<sde::VanDerPolFma, double>, sde::kSaveEndpoint>(c);
  else if (sys == 1) run_fixed<sde::VanDerPolFma, float, sde::Tsit5FastMethod<sde::VanDerPolFma, float>, sde::kSaveEndpoint>(c);
  else return -1;
  return 0;
}

// ---- SimpleEM (csrc/device/sde_em.cuh: em_body, Philox4x32-10 + Box-Muller) ---------------------------------------
namespace {
template <class Sys, class T>
int run_em(int save, int noise_mode, const sde::EMArgs<T>& a) {
  using namespace sde;
  launch_blocks((a.n_traj + g_lanes - 1) / g_lanes, [&]() {
    if (save == kSaveEndpoint) {
      if (noise_mode == kNoisePhilox) em_body<Sys, T, kSaveEndpoint, kNoisePhilox>(a);
      else em_body<Sys, T, kSaveEndpoint, kNoiseProvided>(a);
    } else {
      if (noise_mode == kNoisePhilox) em_body<Sys, T, kSaveEveryStep, kNoisePhilox>(a);
      else em_body<Sys, T, kSaveEveryStep, kNoiseProvided>(a);
    }
  });
  return 0;
}
template <class T>
int em_dispatch(int sys, int save, int layout, int noise_mode, long long n, const void* u0, const void* p, double t0,
                double dt, long long n_steps, unsigned long long seed, long long traj_offset, const void* noise, void* out) {
  sde::EMArgs<T> a;
  std::memset(&a, 0, sizeof a);
  a.u0 = (const T*)u0; a.p = (const T*)p; a.n_traj = n; a.ld_in = n; a.t0 = (T)t0; a.dt = (T)dt; a.n_steps = n_steps;
  a.layout = layout; a.out_u = (T*)out; a.ld_out = n; a.seed = seed; a.traj_offset = traj_offset;
  a.noise = (const T*)noise; a.noise_ld = n;
  switch (sys) {      // registry order of the library: gbm, linadd1, linadd2, ou, nondiag2x4
    case 0: return run_em<sde::EmGBM, T>(save, noise_mode, a);
    case 1: return run_em<sde::EmLinAdd1, T>(save, noise_mode, a);
    case 2: return run_em<sde::EmLinAdd2, T>(save, noise_mode, a);
    case 3: return run_em<sde::EmOU, T>(save, noise_mode, a);
    case 4: return run_em<sde::EmNonDiag2x4, T>(save, noise_mode, a);
  }
  return -1;
}
}  // namespace

extern "C" int emul_em_solve(int sys, int dtype, int save, int layout, int noise_mode, long long n, const void* u0,
                             const void* p, double t0, double dt, long long n_steps, unsigned long long seed,
                             long long traj_offset, const void* noise, void* out) {
  return dtype == 0 ? em_dispatch<double>(sys, save, layout, noise_mode, n, u0, p, t0, dt, n_steps, seed, traj_offset, noise, out)
                    : em_dispatch<float>(sys, save, layout, noise_mode, n, u0, p, t0, dt, n_steps, seed, traj_offset, noise, out);
}

// the normals a Philox solve consumes: out[(step*M + m) * n_traj + traj]
extern "C" int emul_em_noise(int dtype, unsigned long long seed, long long traj_offset, long long n_traj,
                             long long n_normals, void* out) {
  launch_blocks((n_traj + g_lanes - 1) / g_lanes, [&]() {
    if (dtype == 0) sde::em_noise_body<double>(seed, traj_offset, n_traj, n_normals, (double*)out, n_traj);
    else sde::em_noise_body<float>(seed, traj_offset, n_traj, n_normals, (float*)out, n_traj);
  });
  return 0;
}

// 1 = single-lane warps (default), 32 = full warps of 32 host threads (see the header of this file)
extern "C" int emul_set_lanes(int lanes) {
  if (lanes != 1 && lanes != 32) return -1;
  if (g_lanes > 1) pthread_barrier_destroy(&g_bar);
  g_lanes = lanes;
  if (lanes > 1) pthread_barrier_init(&g_bar, nullptr, (unsigned)lanes);
  return 0;
}
