// sde_em.cuh -- ensemble Euler-Maruyama (the reference's SimpleEM, src/euler_maruyama.jl:48-94,
// out-of-place method) for sm_100a: one trajectory per thread, state in registers, every state of a
// trajectory optionally stored (the reference keeps all of them, :67).
//
// Noise.  The reference draws `randn(typeof(u0))` from Julia's task-local RNG (:77,:80,:84), which no
// other process can reproduce.  Here the increments come from a COUNTER-BASED generator, so that a
// trajectory's noise depends only on (seed, global trajectory index, step, component) -- never on the
// launch geometry, the device count or how the ensemble was cut into pieces:
//   the normal with linear index q = step*M + m of global trajectory g is element q % K of Philox4x32-10
//   block b = q / K (K = 2 for double, 4 for float), counter (g_lo, g_hi, b_lo, b_hi), key (seed_lo, seed_hi);
//   double: a = r0 | r1<<32, c = r2 | r3<<32, u1 = ((a>>11)+1) 2^-53 in (0,1], u2 = (c>>11) 2^-53 in [0,1),
//           rad = sqrt(-2 ln u1), z0 = rad cos(2 pi u2), z1 = rad sin(2 pi u2)          (Box-Muller;
//           ln and sin/cos are table-driven FP64 routines accurate to ~2e-16, see NormalBlock<double>)
//   float : (r0,r1) and (r2,r3): u1 = ((r>>8)+1) 2^-24, u2 = (r'>>8) 2^-24, same transform
// or (kNoiseProvided) from a caller-supplied array noise[(step*M + m) * noise_ld + traj], which is how
// the parity tests feed the kernel and the CPU oracle the same increments.
//
// Step arithmetic = the reference's @muladd rewriting (oracle/oracle_em.cpp states the derivation):
//   scalar            u = fma(sqdt*g, z, fma(f, dt, uprev))                       (:76-77)
//   vector diagonal   u_c = fma(f_c, dt, uprev_c + (sqdt*g_c)*z_c)                (:79-80)
//   non-diagonal      u_i = (sum_j (sqdt*G_ij)*z_j, left to right) + fma(f_i, dt, uprev_i)   (:83-84, A11)
//   tprev = fma(i, dt, t0) (:68), sqdt = sqrt(dt) (:69)
#pragma once
#include "sde_common.cuh"
#include "sde_series.cuh"

namespace sde {

enum NoiseMode { kNoisePhilox = 0, kNoiseProvided = 1 };

template <class T>
struct EMArgs {
  const T* u0;        // SoA u0[c * ld_in + i]
  const T* p;         // SoA p [c * ld_in + i]
  i64 n_traj;         // trajectories of this launch
  i64 ld_in;
  T t0, dt;
  i64 n_steps;        // states per trajectory = n_steps + 1
  int layout;
  T* out_u;           // endpoint: SoA out_u[c * ld_out + i]; every step: n_steps + 1 slots, layout as KArgs
  i64 ld_out;
  u64 seed;
  i64 traj_offset;    // global index of trajectory 0 of this launch (Philox counter)
  const T* noise;     // kNoiseProvided: noise[(step*M + m) * noise_ld + i]
  i64 noise_ld;
};

// ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11) ------------------------------------------
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                              unsigned k0, unsigned k1, unsigned* r) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const unsigned hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const unsigned hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1;
    c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}

template <class T> struct NormalBlock;
template <> struct NormalBlock<double> {
  static constexpr int K = 2;
  __device__ __forceinline__ static void make(u64 seed, u64 traj, u64 b, double* z, CtrlTab tab) {
    unsigned r[4];
    philox4x32_10((unsigned)traj, (unsigned)(traj >> 32), (unsigned)b, (unsigned)(b >> 32),
                  (unsigned)seed, (unsigned)(seed >> 32), r);
    const u64 a = (u64)r[0] | ((u64)r[1] << 32), c = (u64)r[2] | ((u64)r[3] << 32);
    const double u1 = (double)((a >> 11) + 1ull) * 1.1102230246251565e-16;   // 2^-53: (0,1]
    const double v = (double)(c >> 11) * 4.4408920985006262e-16;             // 2^-51: 4 u2 in [0,4)
    // rad = sqrt(-2 ln u1), ln through the controller's table-driven log2 (|.|: log2(1) is +2e-18, not 0)
    const double rad = sqrt(fabs(sde_log2_fast(u1, tab) * ctrl_const(tab, kC_neg2ln2)));
    double sn, cs;
    sde_sincos_halfpi(v, tab, &sn, &cs);      // angle 2 pi u2 = (pi/2) v
    z[0] = rad * cs;
    z[1] = rad * sn;
  }
};
template <> struct NormalBlock<float> {
  static constexpr int K = 4;
  __device__ __forceinline__ static void make(u64 seed, u64 traj, u64 b, float* z, CtrlTab) {
    unsigned r[4];
    philox4x32_10((unsigned)traj, (unsigned)(traj >> 32), (unsigned)b, (unsigned)(b >> 32),
                  (unsigned)seed, (unsigned)(seed >> 32), r);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u1 = (float)((r[2 * h] >> 8) + 1u) * 5.9604644775390625e-8f;   // 2^-24, (0,1]
      const float u2 = (float)(r[2 * h + 1] >> 8) * 5.9604644775390625e-8f;      // [0,1)
      const float rad = sqrtf(-2.0f * logf(u1));
      float sn, cs;
      sincospif(2.0f * u2, &sn, &cs);
      z[2 * h] = rad * cs;
      z[2 * h + 1] = rad * sn;
    }
  }
};

// the generator's constants: one shared-memory copy of k_ctrl per CTA (all threads must call this)
__device__ __forceinline__ CtrlTab em_load_table() {
  __shared__ __align__(16) double s_tab[kC_count];
  for (int i = threadIdx.x; i < kC_count; i += blockDim.x) s_tab[i] = k_ctrl[i];
  __syncthreads();
  return s_tab;
}

// sequential reader of one trajectory's normal stream (linear index q = step*M + m)
template <class T, int NOISE>
struct NoiseStream {
  static constexpr int K = NormalBlock<T>::K;
  const EMArgs<T>& a;
  i64 traj;      // local index
  i64 q;
  CtrlTab tab;
  T z[K];
  __device__ __forceinline__ NoiseStream(const EMArgs<T>& a_, i64 traj_, CtrlTab tab_) : a(a_), traj(traj_), q(0), tab(tab_) {}
  __device__ __forceinline__ T next() {
    T v;
    if (NOISE == kNoiseProvided) {
      v = a.noise[q * a.noise_ld + traj];
    } else {
      const int k = (int)(q % K);
      if (k == 0) NormalBlock<T>::make(a.seed, (u64)(a.traj_offset + traj), (u64)(q / K), z, tab);
      v = z[0];
#pragma unroll
      for (int j = 1; j < K; ++j) if (k == j) v = z[j];
    }
    ++q;
    return v;
  }
};

// SAVE: kSaveEndpoint | kSaveEveryStep (the reference's behaviour)
// STAGED: every state of a trajectory goes through the fixed-step kernels' series writer (sde_series.cuh): trajectory-major
// rows staged in shared memory and written in whole 128-byte lines (GBM, 4 Mi paths x 255 steps, FP64: 13.3 -> see
// profiles/r2_em_everystep.txt); the lanes of a warp stay together then (lanes beyond the ensemble integrate a copy of the
// last trajectory and never write)
template <class Sys, class T, int SAVE, int NOISE, bool STAGED>
__device__ __forceinline__ void em_body_impl(const EMArgs<T>& a) {
  constexpr int N = Sys::N, NP = Sys::NP, M = Sys::M;
  constexpr bool kDiag = Sys::kDiagonal;
  constexpr bool kStaged = STAGED && SAVE == kSaveEveryStep;
  CtrlTab tab = nullptr;
  if (NOISE == kNoisePhilox && sizeof(T) == 8) tab = em_load_table();
  const i64 traj = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = traj < a.n_traj;
  if (!kStaged) {
    if (!valid) return;
  } else {
    if (traj - (i64)(threadIdx.x & 31u) >= a.n_traj) return;
  }
  const i64 src = valid ? traj : a.n_traj - 1;
  T u[N], p[NP > 0 ? NP : 1];
#pragma unroll
  for (int c = 0; c < N; ++c) u[c] = a.u0[(i64)c * a.ld_in + src];
#pragma unroll
  for (int c = 0; c < NP; ++c) p[c] = a.p[(i64)c * a.ld_in + src];
  const T dt = a.dt;
  const T sqdt = sde_sqrt(dt);
  NoiseStream<T, NOISE> ns(a, src, tab);
  // the writer only looks at the output description of its argument block
  KArgs<T> ka;
  ka.n_traj = a.n_traj;
  ka.layout = a.layout;
  ka.out_u = a.out_u;
  ka.ld_out = a.ld_out;
  ka.n_out = a.n_steps + 1;
  SeriesWriter<T, N, kStaged> w(ka, traj, valid);
  if (SAVE == kSaveEveryStep) w.put(u);
  for (i64 s = 0; s < a.n_steps; ++s) {
    const T tprev = fma((T)s, dt, a.t0);
    T f[N], g[kDiag ? N : N * M], z[M];
    Sys::rhs(f, u, p, tprev);
    Sys::noise(g, u, p, tprev);
#pragma unroll
    for (int m = 0; m < M; ++m) z[m] = ns.next();
    if (kDiag && N == 1) {
      u[0] = fma(sqdt * g[0], z[0], fma(f[0], dt, u[0]));
    } else if (kDiag) {
#pragma unroll
      for (int c = 0; c < N; ++c) {
        const T odd = u[c] + (sqdt * g[c]) * z[c];
        u[c] = fma(f[c], dt, odd);
      }
    } else {
      T un[N];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        T acc = (sqdt * g[i * M]) * z[0];
#pragma unroll
        for (int j = 1; j < M; ++j) acc = acc + (sqdt * g[i * M + j]) * z[j];
        un[i] = acc + fma(f[i], dt, u[i]);
      }
#pragma unroll
      for (int i = 0; i < N; ++i) u[i] = un[i];
    }
    if (SAVE == kSaveEveryStep) w.put(u);
  }
  if (SAVE == kSaveEveryStep) w.finish();
  if (SAVE == kSaveEndpoint) {
#pragma unroll
    for (int c = 0; c < N; ++c) a.out_u[(i64)c * a.ld_out + traj] = u[c];
  }
}

// staged rows when the launch asked for the trajectory-major layout (the launcher then provides the writer's dynamic
// shared memory: sde_em_api.cu) -- a run-time property of the launch, so both forms live in the every-step kernels
template <class Sys, class T, int SAVE, int NOISE>
__device__ __forceinline__ void em_body(const EMArgs<T>& a) {
  if (SAVE == kSaveEveryStep && a.layout == kLayoutTrajMajor && blockDim.x >= 32) em_body_impl<Sys, T, SAVE, NOISE, true>(a);
  else em_body_impl<Sys, T, SAVE, NOISE, false>(a);
}

// writes the normals a kNoisePhilox solve with the same (seed, traj_offset) consumes:
// out[(step*M + m) * ld + traj]
template <class T>
__device__ __forceinline__ void em_noise_body(u64 seed, i64 traj_offset, i64 n_traj, i64 n_normals, T* out, i64 ld) {
  constexpr int K = NormalBlock<T>::K;
  CtrlTab tab = nullptr;
  if (sizeof(T) == 8) tab = em_load_table();
  const i64 traj = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (traj >= n_traj) return;
  T z[K];
  for (i64 b = 0; b * K < n_normals; ++b) {
    NormalBlock<T>::make(seed, (u64)(traj_offset + traj), (u64)b, z, tab);
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (b * K + j < n_normals) out[(b * K + j) * ld + traj] = z[j];
  }
}

// ---- built-in SDE systems (drift rhs + diffusion noise) -------------------------------------------
//   gbm         docstring example src/euler_maruyama.jl:27-28: f = p1*u, g = p2*u (0.1u, 0.2u)
//   linadd1/2   test/simpleem_tests.jl:4-5,16: f = p1*u (2u), g = p2 (1), scalar / SVector{2}
//   ou          Ornstein-Uhlenbeck (not in the reference): f = p1*(p2 - u), g = p3
//   nondiag2x4  test/simpleem_tests.jl:33-47: f = p1 .* u (1.01), G = the test's 2x4 matrix
struct EmGBM {
  static constexpr int N = 1, NP = 2, M = 1;
  static constexpr bool kDiagonal = true;
  template <class T> __device__ __forceinline__ static void rhs(T* f, const T* u, const T* p, T) { f[0] = p[0] * u[0]; }
  template <class T> __device__ __forceinline__ static void noise(T* g, const T* u, const T* p, T) { g[0] = p[1] * u[0]; }
};
struct EmLinAdd1 {
  static constexpr int N = 1, NP = 2, M = 1;
  static constexpr bool kDiagonal = true;
  template <class T> __device__ __forceinline__ static void rhs(T* f, const T* u, const T* p, T) { f[0] = p[0] * u[0]; }
  template <class T> __device__ __forceinline__ static void noise(T* g, const T*, const T* p, T) { g[0] = p[1]; }
};
struct EmLinAdd2 {
  static constexpr int N = 2, NP = 2, M = 2;
  static constexpr bool kDiagonal = true;
  template <class T> __device__ __forceinline__ static void rhs(T* f, const T* u, const T* p, T) {
    f[0] = p[0] * u[0]; f[1] = p[0] * u[1];
  }
  template <class T> __device__ __forceinline__ static void noise(T* g, const T*, const T* p, T) { g[0] = p[1]; g[1] = p[1]; }
};
struct EmOU {
  static constexpr int N = 1, NP = 3, M = 1;
  static constexpr bool kDiagonal = true;
  template <class T> __device__ __forceinline__ static void rhs(T* f, const T* u, const T* p, T) { f[0] = p[0] * (p[1] - u[0]); }
  template <class T> __device__ __forceinline__ static void noise(T* g, const T*, const T* p, T) { g[0] = p[2]; }
};
struct EmNonDiag2x4 {
  static constexpr int N = 2, NP = 1, M = 4;
  static constexpr bool kDiagonal = false;
  template <class T> __device__ __forceinline__ static void rhs(T* f, const T* u, const T* p, T) {
    f[0] = p[0] * u[0]; f[1] = p[0] * u[1];
  }
  template <class T> __device__ __forceinline__ static void noise(T* g, const T* u, const T*, T) {
    g[0] = T(0.3) * u[0]; g[1] = T(0.6) * u[0]; g[2] = T(0.9) * u[0]; g[3] = T(0.12) * u[0];
    g[4] = T(1.2) * u[1]; g[5] = T(0.2) * u[1]; g[6] = T(0.3) * u[1]; g[7] = T(1.8) * u[1];
  }
};

}  // namespace sde
