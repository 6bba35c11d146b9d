// Host emulation of the device header's controller helpers (csrc/device/sde_common.cuh), so that the
// CPU test suite can check the table-driven log2 / exp2 against an arbitrary-precision reference.
// Test infrastructure: the device intrinsics the header uses are restated with memcpy / std::fma.
#include <cmath>
#include <cstring>
#include <cstdint>
#define SDE_HOST_EMULATION 1
#define __device__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __constant__
#define __align__(n) __attribute__((aligned(n)))
struct double2 { double x, y; };
static inline int __double2hiint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo) {
  int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double x; std::memcpy(&x, &b, 8); return x;
}
static inline double __longlong_as_double(long long v) { double x; std::memcpy(&x, &v, 8); return x; }
static inline long long __double_as_longlong(double x) { long long v; std::memcpy(&v, &x, 8); return v; }
static inline float __int_as_float(int v) { float x; std::memcpy(&x, &v, 4); return x; }
static inline int __float_as_int(float v) { int x; std::memcpy(&x, &v, 4); return x; }
using std::fma;
#include "../simplediffeq.jl_b200/csrc/device/sde_common.cuh"

extern "C" {
void emul_log2(const double* x, double* out, long n) { for (long i = 0; i < n; ++i) out[i] = sde::sde_log2_fast(x[i], sde::k_ctrl); }
void emul_exp2(const double* x, double* out, long n) { for (long i = 0; i < n; ++i) out[i] = sde::sde_exp2_fast(x[i], sde::k_ctrl); }
void emul_rcp(const double* x, double* out, long n) { for (long i = 0; i < n; ++i) out[i] = sde::sde_rcp_fast(x[i], 1.0); }
double emul_max_abs_nan2(double a, double b) { return sde::max_abs_nan2(a, b); }
double emul_min_abs_nan1(double a, double b) { return sde::min_abs_nan1(a, b); }
double emul_jl_max(double a, double b) { return sde::jl_max(a, b); }
double emul_jl_min(double a, double b) { return sde::jl_min(a, b); }
void emul_sincos_halfpi(const double* v, double* sn, double* cs, long n) {
  for (long i = 0; i < n; ++i) sde::sde_sincos_halfpi(v[i], sde::k_ctrl, &sn[i], &cs[i]);
}
// y fixed, x varies: the controller's calls (EEst^beta1, qold^beta2)
void emul_pow_glibc(const double* x, double y, double* out, long n) { for (long i = 0; i < n; ++i) out[i] = sde::sde_pow_glibc(x[i], y); }
// the C library's pow on the same arguments, called from the same process (no numpy / SIMD variant in between)
void host_libm_pow(const double* x, double y, double* out, long n) { for (long i = 0; i < n; ++i) out[i] = std::pow(x[i], y); }
void emul_powf_glibc(const float* x, float y, float* out, long n) { for (long i = 0; i < n; ++i) out[i] = sde::sde_powf_glibc(x[i], y); }
void host_libm_powf(const float* x, float y, float* out, long n) { for (long i = 0; i < n; ++i) out[i] = std::pow(x[i], y); }
// element-wise (x[i], y[i]) pairs
void emul_pow_glibc_xy(const double* x, const double* y, double* out, long n) { for (long i = 0; i < n; ++i) out[i] = sde::sde_pow_glibc(x[i], y[i]); }
void host_libm_pow_xy(const double* x, const double* y, double* out, long n) { for (long i = 0; i < n; ++i) out[i] = std::pow(x[i], y[i]); }
void emul_powf_glibc_xy(const float* x, const float* y, float* out, long n) { for (long i = 0; i < n; ++i) out[i] = sde::sde_powf_glibc(x[i], y[i]); }
void host_libm_powf_xy(const float* x, const float* y, float* out, long n) { for (long i = 0; i < n; ++i) out[i] = std::pow(x[i], y[i]); }
int emul_ctrl_count() { return sde::kC_count; }
double emul_ctrl(int i) { return sde::k_ctrl[i]; }
}
