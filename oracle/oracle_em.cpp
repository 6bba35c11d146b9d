// oracle/oracle_em.cpp -- CPU restatement of SimpleDiffEq.jl's SimpleEM (Euler-Maruyama) solve body,
// out-of-place method, src/euler_maruyama.jl:48-94.
//
// *** TEST INFRASTRUCTURE ONLY *** (same rules as oracle.cpp: loaded by tests/ and smoke() only).
//
// PARITY STATUS: step arithmetic pinned to the reference's own SOURCE TEXT; noise necessarily not.
// The reference draws its noise from Julia's task-local default RNG (`randn(typeof(u0))`,
// src/euler_maruyama.jl:76-85), which is not reproducible outside that Julia process, and its own tests
// (test/simpleem_tests.jl) only check `sol.t == collect(0:0.25:1.0)` and `length(sol.u) == 5`.  So:
//   * oracle/jlmini parses and executes src/euler_maruyama.jl:46-94 (its @muladd rewriting included) with
//     `randn` replaced by a supplied list of normals; oracle_em_solve reproduces those outputs BIT FOR
//     BIT for scalar and diagonal SVector states, FP64 and FP32 (tests/golden/golden_jlmini_em_v1.json,
//     tests/test_oracle_em_jlmini.py); the non-diagonal branch is outside that pin (assumption A11);
//   * the STEP ARITHMETIC is restated here with the increments dW handed in as data
//     (oracle_em_solve), which makes it comparable bit for bit with the CUDA kernel given the same
//     normals, and against closed-form results (noise-free limit, exact GBM path, strong order 1/2);
//   * the NOISE SPECIFICATION of the CUDA path (Philox4x32-10 counter layout + Box-Muller)
//     is restated independently in oracle_em_normals with libm log/sincos, pinned by the published
//     Random123 known-answer vectors for Philox4x32-10 (tests/test_oracle_em.py).
//
// @muladd placement (MuladdMacro: non-product summands are added first, then every `*` summand is
// folded in with muladd, left to right; a dotted `.*` inside an undotted `+` is not a product for
// the macro):
//   scalar state      u[i] = uprev + f*dt + sqdt*g*randn()            (:76-77)
//                       -> muladd(sqdt*g, z, muladd(f, dt, uprev))
//   vector, diagonal  u[i] = uprev + f*dt + sqdt*g .* randn(SVector)   (:79-80)
//                       -> muladd(f, dt, uprev + (sqdt*g) .* z)        element-wise
//   non-diagonal      u[i] = uprev + f*dt + sqdt*g*randn(m)            (:83-84)
//                       -> muladd(sqdt*G, z, muladd(f, dt, uprev)); the matrix-vector muladd has no
//                          defined rounding order in the reference (BLAS gemv / StaticArrays
//                          generated code).  ASSUMPTION A11, fixed here and in the kernel:
//                          (sum_j (sqdt*G_ij)*z_j, left to right, unfused) + muladd(f_i, dt, uprev_i).
//   time grid         t = [tspan[1] + i*dt for i in 0:n-1] -> muladd(i, dt, tspan[1])   (:68)
//   n = Int((tspan[2]-tspan[1])/dt) + 1 (:66) -- the host layer computes it and raises Julia's
//   InexactError when the quotient is not an integer; sqdt = sqrt(dt) (:69).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ---- SDE systems (ids private to the oracle; formulas stated in DESIGN.md) -------------------
//   gbm        docstring example src/euler_maruyama.jl:27-28   f = p1*u, g = p2*u          (0.1u, 0.2u)
//   linadd1/2  test/simpleem_tests.jl:4-5,16                   f = p1*u, g = p2            (2u, 1), N = 1 / 2
//   ou         (not in the reference) Ornstein-Uhlenbeck        f = p1*(p2 - u), g = p3
//   nondiag2x4 test/simpleem_tests.jl:33-47                    f = p1 .* u, G = 2x4 matrix of the test
enum { EM_GBM = 0, EM_LINADD1 = 1, EM_LINADD2 = 2, EM_OU = 3, EM_NONDIAG2X4 = 4 };

struct Dims { int N, NP, M; bool diagonal; };
bool dims_of(int sys, Dims* d) {
  switch (sys) {
    case EM_GBM: *d = {1, 2, 1, true}; return true;
    case EM_LINADD1: *d = {1, 2, 1, true}; return true;
    case EM_LINADD2: *d = {2, 2, 2, true}; return true;
    case EM_OU: *d = {1, 3, 1, true}; return true;
    case EM_NONDIAG2X4: *d = {2, 1, 4, false}; return true;
  }
  return false;
}

template <class T>
void drift(int sys, T* f, const T* u, const T* p, T) {
  switch (sys) {
    case EM_GBM: f[0] = p[0] * u[0]; break;
    case EM_LINADD1: f[0] = p[0] * u[0]; break;
    case EM_LINADD2: f[0] = p[0] * u[0]; f[1] = p[0] * u[1]; break;
    case EM_OU: f[0] = p[0] * (p[1] - u[0]); break;
    case EM_NONDIAG2X4: f[0] = p[0] * u[0]; f[1] = p[0] * u[1]; break;
  }
}

// g: N entries (diagonal) or N*M row-major
template <class T>
void diffusion(int sys, T* g, const T* u, const T* p, T) {
  switch (sys) {
    case EM_GBM: g[0] = p[1] * u[0]; break;
    case EM_LINADD1: g[0] = p[1]; break;
    case EM_LINADD2: g[0] = p[1]; g[1] = p[1]; break;
    case EM_OU: g[0] = p[2]; break;
    case EM_NONDIAG2X4:
      g[0] = T(0.3) * u[0]; g[1] = T(0.6) * u[0]; g[2] = T(0.9) * u[0]; g[3] = T(0.12) * u[0];
      g[4] = T(1.2) * u[1]; g[5] = T(0.2) * u[1]; g[6] = T(0.3) * u[1]; g[7] = T(1.8) * u[1];
      break;
  }
}

// one trajectory; noise z[(s*M + m) * ld_noise], out[(s) * N + c] for s = 0..n_steps (every state)
template <class T>
void em_trajectory(int sys, const Dims& d, const T* u0, const T* p, T t0, T dt, int64_t n_steps,
                   const T* z, int64_t ld_noise, T* out /* (n_steps+1) x N */) {
  const int N = d.N, M = d.M;
  T u[4], f[4], g[16];
  for (int c = 0; c < N; ++c) { u[c] = u0[c]; out[c] = u0[c]; }
  const T sqdt = std::sqrt(dt);                                        // :69
  for (int64_t s = 0; s < n_steps; ++s) {
    const T tprev = std::fma((T)s, dt, t0);                            // t[i-1], :68
    drift<T>(sys, f, u, p, tprev);
    diffusion<T>(sys, g, u, p, tprev);
    const T* zs = z + (size_t)s * M * ld_noise;
    if (d.diagonal && N == 1) {                                        // :76-77
      u[0] = std::fma(sqdt * g[0], zs[0], std::fma(f[0], dt, u[0]));
    } else if (d.diagonal) {                                           // :79-80
      for (int c = 0; c < N; ++c) {
        const T x = (sqdt * g[c]) * zs[(size_t)c * ld_noise];
        const T odd = u[c] + x;
        u[c] = std::fma(f[c], dt, odd);
      }
    } else {                                                           // :83-84, assumption A11
      T un[4];
      for (int i = 0; i < N; ++i) {
        T acc = (sqdt * g[i * M + 0]) * zs[0];
        for (int j = 1; j < M; ++j) acc = acc + (sqdt * g[i * M + j]) * zs[(size_t)j * ld_noise];
        un[i] = acc + std::fma(f[i], dt, u[i]);
      }
      for (int i = 0; i < N; ++i) u[i] = un[i];
    }
    for (int c = 0; c < N; ++c) out[(size_t)(s + 1) * N + c] = u[c];
  }
}

// ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11)
inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += W0; k1 += W1;
  }
}

// Noise specification: the normal with linear index q = step*M + m of trajectory g
// (GLOBAL index) comes from Philox block b = q / K, K = 2 (f64) or 4 (f32), counter
// (g_lo, g_hi, b_lo, b_hi), key (seed_lo, seed_hi):
//   f64: a = r0 | r1<<32, b = r2 | r3<<32, u1 = ((a>>11)+1) 2^-53 in (0,1], u2 = (b>>11) 2^-53 in [0,1)
//        rad = sqrt(-2 ln u1), z[0] = rad cos(2 pi u2), z[1] = rad sin(2 pi u2)
//   f32: pairs (r0,r1) and (r2,r3): u1 = ((r>>8)+1) 2^-24, u2 = (r'>>8) 2^-24, same transform in float
template <class T> struct Gen;
template <> struct Gen<double> {
  static constexpr int K = 2;
  static void block(uint64_t seed, uint64_t traj, uint64_t b, double* z) {
    uint32_t c[4] = {(uint32_t)traj, (uint32_t)(traj >> 32), (uint32_t)b, (uint32_t)(b >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = (uint64_t)c[0] | ((uint64_t)c[1] << 32), bb = (uint64_t)c[2] | ((uint64_t)c[3] << 32);
    const double u1 = (double)((a >> 11) + 1) * 0x1p-53, u2 = (double)(bb >> 11) * 0x1p-53;
    const double rad = std::sqrt(-2.0 * std::log(u1));
    const double ang = 6.283185307179586476925286766559 * u2;
    z[0] = rad * std::cos(ang);
    z[1] = rad * std::sin(ang);
  }
};
template <> struct Gen<float> {
  static constexpr int K = 4;
  static void block(uint64_t seed, uint64_t traj, uint64_t b, float* z) {
    uint32_t c[4] = {(uint32_t)traj, (uint32_t)(traj >> 32), (uint32_t)b, (uint32_t)(b >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    for (int h = 0; h < 2; ++h) {
      const double u1 = (double)((c[2 * h] >> 8) + 1) * 0x1p-24, u2 = (double)(c[2 * h + 1] >> 8) * 0x1p-24;
      const double rad = std::sqrt(-2.0 * std::log(u1));      // evaluated in double, rounded once
      const double ang = 6.283185307179586476925286766559 * u2;
      z[2 * h] = (float)(rad * std::cos(ang));
      z[2 * h + 1] = (float)(rad * std::sin(ang));
    }
  }
};

template <class T>
void fill_normals(uint64_t seed, int64_t traj_offset, int64_t n_traj, int64_t n_steps, int M, T* out, int64_t ld) {
  constexpr int K = Gen<T>::K;
  const int64_t total = n_steps * M;
  for (int64_t i = 0; i < n_traj; ++i) {
    T z[K];
    for (int64_t q = 0; q < total; ++q) {
      if (q % K == 0) Gen<T>::block(seed, (uint64_t)(traj_offset + i), (uint64_t)(q / K), z);
      out[(size_t)q * ld + i] = z[q % K];
    }
  }
}

template <class T>
int em_solve_t(int sys, int64_t n_traj, const T* u0, const T* p, double t0, double dt, int64_t n_steps,
               const T* noise, T* out, int n_threads) {
  Dims d;
  if (!dims_of(sys, &d)) return -1;
  const int N = d.N, NP = d.NP;
  auto work = [&](int64_t lo, int64_t hi) {
    std::vector<T> row((size_t)(n_steps + 1) * N);
    for (int64_t i = lo; i < hi; ++i) {
      T u[4], pp[4];
      for (int c = 0; c < N; ++c) u[c] = u0[(size_t)c * n_traj + i];
      for (int c = 0; c < NP; ++c) pp[c] = p[(size_t)c * n_traj + i];
      em_trajectory<T>(sys, d, u, pp, (T)t0, (T)dt, n_steps, noise + i, n_traj, row.data());
      std::memcpy(out + (size_t)i * (n_steps + 1) * N, row.data(), row.size() * sizeof(T));
    }
  };
  if (n_threads < 1) n_threads = 1;
  std::vector<std::thread> th;
  for (int k = 0; k < n_threads; ++k) {
    const int64_t lo = n_traj * k / n_threads, hi = n_traj * (k + 1) / n_threads;
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto& t : th) t.join();
  return 0;
}

}  // namespace

extern "C" {

int oracle_em_dims(int sys, int* n_state, int* n_param, int* n_noise, int* diagonal) {
  Dims d;
  if (!dims_of(sys, &d)) return -1;
  *n_state = d.N; *n_param = d.NP; *n_noise = d.M; *diagonal = d.diagonal ? 1 : 0;
  return 0;
}

// u0 [N][n_traj], p [NP][n_traj] (SoA); noise [n_steps][M][n_traj] standard normals;
// out [n_traj][n_steps+1][N] (trajectory major, every state incl. u0).  dtype 0 = f64, 1 = f32.
int oracle_em_solve(int sys, int dtype, int64_t n_traj, const void* u0, const void* p, double t0, double dt,
                    int64_t n_steps, const void* noise, void* out, int n_threads) {
  if (dtype == 0)
    return em_solve_t<double>(sys, n_traj, (const double*)u0, (const double*)p, t0, dt, n_steps,
                              (const double*)noise, (double*)out, n_threads);
  return em_solve_t<float>(sys, n_traj, (const float*)u0, (const float*)p, t0, dt, n_steps,
                           (const float*)noise, (float*)out, n_threads);
}

// the normals of the CUDA path's noise specification: out [n_steps][M][n_traj]
int oracle_em_normals(int dtype, uint64_t seed, int64_t traj_offset, int64_t n_traj, int64_t n_steps, int M,
                      void* out) {
  if (dtype == 0) fill_normals<double>(seed, traj_offset, n_traj, n_steps, M, (double*)out, n_traj);
  else fill_normals<float>(seed, traj_offset, n_traj, n_steps, M, (float*)out, n_traj);
  return 0;
}

// raw Philox4x32-10 block, for the Random123 known-answer vectors
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  philox4x32_10(c, key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}

}  // extern "C"
