// sde_systems.cuh -- built-in right-hand-side registry (device functions).
//
// The RHS arithmetic is written exactly as the (unfused) Julia expression it stands for and the
// translation units are compiled with -fmad=false, so a*b+c here is a rounded product followed by a
// rounded sum -- the reference never applies @muladd to the user's f.
//
//   lorenz        test/gpusimpleatsit5_tests.jl:3-13 of the reference
//   lineardecay   test/gpu_ode_regression.jl:2-4      f(u,p,t) = -u
//   scalargrowth  docstring example src/tsit5/gpuatsit5.jl:33, f(u,p,t) = p1*u (p1 = 1.01)
//   vanderpol / robertson / nbody / nonautonomous have no definition in the reference; their
//   formulas (operation order included) are fixed here and in DESIGN.md.
#pragma once
#include "sde_common.cuh"

namespace sde {

struct Lorenz {
  static constexpr int N = 3, NP = 3;
  template <class T>
  __device__ __forceinline__ static void rhs(T* du, const T* u, const T* p, T) {
    du[0] = p[0] * (u[1] - u[0]);
    du[1] = u[0] * (p[1] - u[2]) - u[1];
    du[2] = u[0] * u[1] - p[2] * u[2];
  }
};

// SDE_COMPAT_FAST_RHS twin of Lorenz: the same right-hand side with its multiply-add pairs contracted (8 -> 6 FP64
// instructions per evaluation; a fixed-step Tsit5 step 126 -> 114).  NOT the reference's arithmetic (its f is never
// under @muladd): results differ from the reference-exact kernels in the last bits of every stage -- <= 1e-12 relative
// over BASELINE config 2's non-chaotic rho-sweep except next to the homoclinic bifurcation at rho = 13.926 (6e-12; GPU
// test), unbounded in a chaotic regime like any rounding change.
struct LorenzFma {
  static constexpr int N = 3, NP = 3;
  template <class T>
  __device__ __forceinline__ static void rhs(T* du, const T* u, const T* p, T) {
    du[0] = p[0] * (u[1] - u[0]);
    du[1] = fma(u[0], p[1] - u[2], -u[1]);
    du[2] = fma(u[0], u[1], -(p[2] * u[2]));
  }
};

// u1' = u2 ; u2' = p1*(1-u1*u1)*u2 - u1      (Julia: p[1]*(1-u[1]*u[1])*u[2]-u[1])
struct VanDerPol {
  static constexpr int N = 2, NP = 1;
  template <class T>
  __device__ __forceinline__ static void rhs(T* du, const T* u, const T* p, T) {
    du[0] = u[1];
    du[1] = (p[0] * (T(1) - u[0] * u[0])) * u[1] - u[0];
  }
};

// SDE_COMPAT_FAST_RHS twin of VanDerPol (5 -> 3 FP64 instructions per evaluation)
struct VanDerPolFma {
  static constexpr int N = 2, NP = 1;
  template <class T>
  __device__ __forceinline__ static void rhs(T* du, const T* u, const T* p, T) {
    du[0] = u[1];
    du[1] = fma(p[0] * fma(-u[0], u[0], T(1)), u[1], -u[0]);
  }
};

// Robertson kinetics with caller-chosen (non-stiff) rates p1,p2,p3:
//   u1' = -p1*u1 + p3*u2*u3 ; u2' = p1*u1 - p2*u2*u2 - p3*u2*u3 ; u3' = p2*u2*u2
struct Robertson {
  static constexpr int N = 3, NP = 3;
  template <class T>
  __device__ __forceinline__ static void rhs(T* du, const T* u, const T* p, T) {
    du[0] = (-p[0]) * u[0] + (p[2] * u[1]) * u[2];
    du[1] = (p[0] * u[0] - (p[1] * u[1]) * u[1]) - (p[2] * u[1]) * u[2];
    du[2] = (p[1] * u[1]) * u[1];
  }
};

// "N-body-lite": 3 planar bodies, G = 1, masses p1..p3, Plummer softening eps^2 = 1e-4.
// state (x1,y1,x2,y2,x3,y3, vx1,vy1,vx2,vy2,vx3,vy3)
struct NBodyLite {
  static constexpr int N = 12, NP = 3;
  template <class T>
  __device__ __forceinline__ static void rhs(T* du, const T* u, const T* p, T) {
#pragma unroll
    for (int i = 0; i < 6; ++i) du[i] = u[6 + i];
    const T eps2 = T(1.0e-4);
    T ax[3] = {T(0), T(0), T(0)}, ay[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        if (a == b) continue;
        T dx = u[2 * b] - u[2 * a];
        T dy = u[2 * b + 1] - u[2 * a + 1];
        T r2 = (dx * dx + dy * dy) + eps2;
        T w = p[b] / (r2 * sde_sqrt(r2));
        ax[a] = ax[a] + w * dx;
        ay[a] = ay[a] + w * dy;
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      du[6 + 2 * a] = ax[a];
      du[7 + 2 * a] = ay[a];
    }
  }
};

struct LinearDecay {
  static constexpr int N = 3, NP = 3;
  template <class T>
  __device__ __forceinline__ static void rhs(T* du, const T* u, const T*, T) {
#pragma unroll
    for (int i = 0; i < 3; ++i) du[i] = -u[i];
  }
};

struct ScalarGrowth {
  static constexpr int N = 1, NP = 1;
  template <class T>
  __device__ __forceinline__ static void rhs(T* du, const T* u, const T* p, T) {
    du[0] = p[0] * u[0];
  }
};

// time-dependent test system (exercises the stage-time arguments, quirks Q1/Q3):
//   u1' = u2 + t ; u2' = -p1*u1 + p2*t*t
struct NonAutonomous {
  static constexpr int N = 2, NP = 2;
  template <class T>
  __device__ __forceinline__ static void rhs(T* du, const T* u, const T* p, T t) {
    du[0] = u[1] + t;
    du[1] = (-p[0]) * u[0] + (p[1] * t) * t;
  }
};

}  // namespace sde
