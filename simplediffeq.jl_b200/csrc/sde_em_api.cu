// sde_em_api.cu -- C ABI of the ensemble Euler-Maruyama path (the reference's SimpleEM,
// src/euler_maruyama.jl:48-94).  Kernels: device/sde_em.cuh.
#include <cuda_runtime.h>
#include <nvrtc.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/simplediffeq_cuda.h"
#include "device/sde_em.cuh"
#include "sde_internal.h"

using namespace sde_host;

static_assert((int)sde::kNoisePhilox == SDE_NOISE_PHILOX && (int)sde::kNoiseProvided == SDE_NOISE_PROVIDED, "noise ids");

namespace sde {

template <class Sys, class T, int SAVE, int NOISE>
__global__ void __launch_bounds__(sde_host::kBlock) em_kernel(const __grid_constant__ EMArgs<T> a) {
  em_body<Sys, T, SAVE, NOISE>(a);
}

template <class T>
__global__ void __launch_bounds__(sde_host::kBlock) em_noise_kernel(u64 seed, i64 traj_offset, i64 n_traj, i64 n_normals,
                                                                    T* out, i64 ld) {
  em_noise_body<T>(seed, traj_offset, n_traj, n_normals, out, ld);
}

template <class Sys, class T>
const void* em_lookup_t(int save, int noise) {
  if (save == kSaveEndpoint)
    return noise == kNoisePhilox ? (const void*)&em_kernel<Sys, T, kSaveEndpoint, kNoisePhilox>
                                 : (const void*)&em_kernel<Sys, T, kSaveEndpoint, kNoiseProvided>;
  return noise == kNoisePhilox ? (const void*)&em_kernel<Sys, T, kSaveEveryStep, kNoisePhilox>
                               : (const void*)&em_kernel<Sys, T, kSaveEveryStep, kNoiseProvided>;
}
template <class Sys>
const void* em_lookup(int dtype, int save, int noise) {
  return dtype == SDE_F64 ? em_lookup_t<Sys, double>(save, noise) : em_lookup_t<Sys, float>(save, noise);
}

}  // namespace sde

namespace {

struct EmCompiled {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kernel = nullptr;
  std::vector<char> cubin;
};

}  // namespace

struct sde_em_system_s {
  bool builtin = true;
  std::string name;
  int n_state = 0, n_param = 0, n_noise = 0;
  bool diagonal = true;
  const void* (*lookup)(int dtype, int save, int noise) = nullptr;
  std::string src;
  std::mutex mu;
  std::map<std::string, EmCompiled> cache;
};

namespace {

struct EmBuiltin {
  const char* name;
  int n_state, n_param, n_noise;
  bool diagonal;
  const void* (*lookup)(int, int, int);
};
const EmBuiltin kEmBuiltins[] = {
    {"gbm", 1, 2, 1, true, sde::em_lookup<sde::EmGBM>},
    {"linadd1", 1, 2, 1, true, sde::em_lookup<sde::EmLinAdd1>},
    {"linadd2", 2, 2, 2, true, sde::em_lookup<sde::EmLinAdd2>},
    {"ou", 1, 3, 1, true, sde::em_lookup<sde::EmOU>},
    {"nondiag2x4", 2, 1, 4, false, sde::em_lookup<sde::EmNonDiag2x4>},
};
constexpr int kNumEmBuiltins = (int)(sizeof(kEmBuiltins) / sizeof(kEmBuiltins[0]));
sde_em_system_s g_em_handles[kNumEmBuiltins];
std::once_flag g_em_once;

void init_em_builtins() {
  for (int i = 0; i < kNumEmBuiltins; ++i) {
    sde_em_system_s& h = g_em_handles[i];
    h.builtin = true;
    h.name = kEmBuiltins[i].name;
    h.n_state = kEmBuiltins[i].n_state;
    h.n_param = kEmBuiltins[i].n_param;
    h.n_noise = kEmBuiltins[i].n_noise;
    h.diagonal = kEmBuiltins[i].diagonal;
    h.lookup = kEmBuiltins[i].lookup;
  }
}

int em_validate(const sde_em_system_s* sys, const sde_em_options_t* o) {
  if (!sys || !o) return fail(SDE_ERR_INVALID, "null system or options");
  if (o->dtype != SDE_F64 && o->dtype != SDE_F32) return fail(SDE_ERR_INVALID, "unknown dtype %d", o->dtype);
  if (o->save_mode != SDE_SAVE_ENDPOINT && o->save_mode != SDE_SAVE_EVERYSTEP)
    return fail(SDE_ERR_UNSUPPORTED, "SimpleEM saves every step (the reference) or the endpoint only; save_mode %d", o->save_mode);
  if (o->layout != SDE_LAYOUT_TRAJ_MAJOR && o->layout != SDE_LAYOUT_SOA) return fail(SDE_ERR_INVALID, "unknown layout %d", o->layout);
  if (o->noise_mode != SDE_NOISE_PHILOX && o->noise_mode != SDE_NOISE_PROVIDED)
    return fail(SDE_ERR_INVALID, "unknown noise_mode %d", o->noise_mode);
  if (o->n_traj < 0 || o->n_steps < 0) return fail(SDE_ERR_INVALID, "n_traj / n_steps < 0");
  if (o->traj_offset < 0) return fail(SDE_ERR_INVALID, "traj_offset < 0");
  return SDE_OK;
}

std::string em_user_program(const sde_em_system_s* sys, int dtype, int save, int noise, bool syntax_only) {
  std::string s;
  s += dtype == SDE_F64 ? "typedef double real;\n" : "typedef float real;\n";
  s += "#include \"sde_em.cuh\"\n#line 1 \"user_sde.cu\"\n";
  s += sys->src;
  s += "\n#line 1 \"sde_user_em_glue.cu\"\n";
  char buf[1024];
  snprintf(buf, sizeof buf,
           "struct SdeUserEm {\n"
           "  static constexpr int N = %d, NP = %d, M = %d;\n"
           "  static constexpr bool kDiagonal = %s;\n"
           "  template <class T> __device__ __forceinline__ static void rhs(T* f, const T* u, const T* p, T t) { ::rhs(f, u, p, t); }\n"
           "  template <class T> __device__ __forceinline__ static void noise(T* g, const T* u, const T* p, T t) { ::noise(g, u, p, t); }\n"
           "};\n",
           sys->n_state, sys->n_param, sys->n_noise, sys->diagonal ? "true" : "false");
  s += buf;
  if (syntax_only) {
    s += "extern \"C\" __global__ void sde_user_em_check(real* f, real* g, const real* u, const real* p, real t) {\n"
         "  SdeUserEm::rhs<real>(f, u, p, t);\n  SdeUserEm::noise<real>(g, u, p, t);\n}\n";
    return s;
  }
  snprintf(buf, sizeof buf,
           "extern \"C\" __global__ void __launch_bounds__(%d) sde_user_em_kernel(const __grid_constant__ sde::EMArgs<real> a) {\n"
           "  sde::em_body<SdeUserEm, real, %d, %d>(a);\n}\n",
           kBlock, save, noise);
  s += buf;
  return s;
}

int em_get_kernel(sde_em_system_s* sys, const sde_em_options_t* o, bool load, const void** fn) {
  if (sys->builtin) {
    *fn = sys->lookup(o->dtype, o->save_mode, o->noise_mode);
    return *fn ? SDE_OK : fail(SDE_ERR_UNSUPPORTED, "no SimpleEM kernel for system %s", sys->name.c_str());
  }
  int dev = -1;
  if (load) SDE_CUDA(cudaGetDevice(&dev));
  char key[64];
  snprintf(key, sizeof key, "%d/%d/%d", o->dtype, o->save_mode, o->noise_mode);
  std::lock_guard<std::mutex> lk(sys->mu);
  EmCompiled& c = sys->cache[key];
  if (c.cubin.empty()) {
    int rc = nvrtc_compile(em_user_program(sys, o->dtype, o->save_mode, o->noise_mode, false), &c.cubin, nullptr);
    if (rc != SDE_OK) { sys->cache.erase(key); return rc; }
  }
  if (!load) return SDE_OK;
  char dkey[80];
  snprintf(dkey, sizeof dkey, "%s@%d", key, dev);
  EmCompiled& d = sys->cache[dkey];
  if (!d.kernel) {
    const std::vector<char>& cubin = sys->cache[key].cubin;
    SDE_CUDA(cudaLibraryLoadData(&d.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    SDE_CUDA(cudaLibraryGetKernel(&d.kernel, d.lib, "sde_user_em_kernel"));
  }
  *fn = (const void*)d.kernel;
  return SDE_OK;
}

int64_t em_slots(const sde_em_options_t* o) { return o->save_mode == SDE_SAVE_EVERYSTEP ? o->n_steps + 1 : 1; }

template <class T>
int em_launch_t(sde_em_system_s* sys, const sde_em_options_t* o, const void* fn, const void* d_u0, const void* d_p,
                int64_t ld_in, const void* d_noise, int64_t noise_ld, void* d_out, int64_t ld_out, cudaStream_t st) {
  if (o->n_traj <= 0) return SDE_OK;
  sde::EMArgs<T> a;
  memset(&a, 0, sizeof a);
  a.u0 = (const T*)d_u0;
  a.p = (const T*)d_p;
  a.n_traj = o->n_traj;
  a.ld_in = ld_in;
  a.t0 = (T)o->t0;
  a.dt = (T)o->dt;
  a.n_steps = o->n_steps;
  a.layout = o->layout;
  a.out_u = (T*)d_out;
  a.ld_out = ld_out;
  a.seed = o->seed;
  a.traj_offset = o->traj_offset;
  a.noise = (const T*)d_noise;
  a.noise_ld = noise_ld;
  const int64_t grid = (o->n_traj + kBlock - 1) / kBlock;
  if (grid > 0x7fffffffLL) return fail(SDE_ERR_INVALID, "n_traj too large for one launch");
  // every state kept, trajectory-major rows: the kernel takes the staged series writer (sde_em.cuh: em_body), which
  // needs its dynamic shared memory
  size_t smem = 0;
  if (o->save_mode == SDE_SAVE_EVERYSTEP && o->layout == SDE_LAYOUT_TRAJ_MAJOR) {
    smem = staged_smem_bytes(sys->n_state, sizeof(T), kBlock, false);
    if (smem > 48 * 1024)
      SDE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  void* params[] = {&a};
  SDE_CUDA(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(kBlock), params, smem, st));
  g_launches.fetch_add(1);
  return SDE_OK;
}

int em_launch(sde_em_system_s* sys, const sde_em_options_t* o, const void* fn, const void* d_u0, const void* d_p,
              int64_t ld_in, const void* d_noise, int64_t noise_ld, void* d_out, int64_t ld_out, cudaStream_t st) {
  if (o->noise_mode == SDE_NOISE_PROVIDED && !d_noise && o->n_steps > 0)
    return fail(SDE_ERR_INVALID, "noise_mode SDE_NOISE_PROVIDED needs a noise array");
  if (o->dtype == SDE_F64) return em_launch_t<double>(sys, o, fn, d_u0, d_p, ld_in, d_noise, noise_ld, d_out, ld_out, st);
  return em_launch_t<float>(sys, o, fn, d_u0, d_p, ld_in, d_noise, noise_ld, d_out, ld_out, st);
}

int em_noise_launch(int dtype, uint64_t seed, int64_t traj_offset, int64_t n_traj, int64_t n_normals, void* d_out,
                    int64_t ld, cudaStream_t st) {
  if (n_traj <= 0 || n_normals <= 0) return SDE_OK;
  const int64_t grid = (n_traj + kBlock - 1) / kBlock;
  if (grid > 0x7fffffffLL) return fail(SDE_ERR_INVALID, "n_traj too large for one launch");
  if (dtype == SDE_F64)
    sde::em_noise_kernel<double><<<(unsigned)grid, kBlock, 0, st>>>(seed, traj_offset, n_traj, n_normals, (double*)d_out, ld);
  else
    sde::em_noise_kernel<float><<<(unsigned)grid, kBlock, 0, st>>>(seed, traj_offset, n_traj, n_normals, (float*)d_out, ld);
  SDE_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return SDE_OK;
}

// one device of a host-buffer solve: trajectories [lo, hi) in pieces that fit the memory budget
int em_solve_shard(sde_em_system_s* sys, const sde_em_options_t* o, int device, int64_t lo, int64_t hi, const char* u0,
                   const char* p, const char* noise, char* out_u, std::string* err) {
  // the caller's current device is restored on every exit path (a single-device solve runs on the caller's thread)
  struct DeviceGuard {
    int prev = -1;
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  } guard;
  auto body = [&]() -> int {
    if (device >= 0) {
      int cur = -1;
      SDE_CUDA(cudaGetDevice(&cur));
      if (cur != device) { SDE_CUDA(cudaSetDevice(device)); guard.prev = cur; }
    }
    int dev = 0;
    SDE_CUDA(cudaGetDevice(&dev));
    if (hi <= lo) return SDE_OK;
    const size_t es = esize(o->dtype);
    const int N = sys->n_state, NP = sys->n_param, M = sys->n_noise;
    const int64_t n_all = o->n_traj, slots = em_slots(o);
    const int64_t n_normals = o->n_steps * M;
    const bool provided = o->noise_mode == SDE_NOISE_PROVIDED;
    const void* fn = nullptr;
    int rc = em_get_kernel(sys, o, true, &fn);
    if (rc != SDE_OK) return rc;
    cudaMemPool_t pool;
    rc = device_pool(dev, &pool);
    if (rc != SDE_OK) return rc;
    const size_t per_traj = es * ((size_t)N + NP + (size_t)N * slots + (provided ? (size_t)n_normals : 0));
    int64_t piece = hi - lo;
    // large solves: pieces bounded by what is free now (cudaMemGetInfo costs ~0.3 ms, so small solves skip it like
    // sde_solve does: any B200 has 512 MB to spare; SDE_TUNE_EM_MEMINFO = measurement only)
    if ((double)per_traj * (double)(hi - lo) > 512.0 * 1048576.0 || getenv("SDE_TUNE_EM_MEMINFO")) {
      size_t free_b = 0, total_b = 0;
      SDE_CUDA(cudaMemGetInfo(&free_b, &total_b));
      piece = (int64_t)std::max<size_t>(1, (size_t)(0.6 * (double)free_b) / per_traj);
    }
    if (const char* e = getenv("SDE_TUNE_PIECE")) piece = std::max<int64_t>(32, atoll(e));   // measurement / tests only
    piece = std::min<int64_t>(piece, hi - lo);
    if (piece > 32) piece = (piece + 31) & ~(int64_t)31;
    cudaStream_t st = nullptr;
    char *d_u0 = nullptr, *d_p = nullptr, *d_z = nullptr, *d_out = nullptr;
    auto cleanup = [&]() {
      if (!st) return;
      cudaStreamSynchronize(st);
      void* ptrs[] = {d_u0, d_p, d_z, d_out};
      for (void* q : ptrs) if (q) cudaFreeAsync(q, st);
      cudaStreamSynchronize(st);
      cudaStreamDestroy(st);
      cudaMemPoolTrimTo(pool, pool_keep_bytes());
    };
#define SDE_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); \
      return fail(SDE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
    SDE_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    SDE_TRY(cudaMallocFromPoolAsync((void**)&d_u0, es * N * piece, pool, st));
    if (NP) SDE_TRY(cudaMallocFromPoolAsync((void**)&d_p, es * NP * piece, pool, st));
    if (provided && n_normals) SDE_TRY(cudaMallocFromPoolAsync((void**)&d_z, es * n_normals * piece, pool, st));
    SDE_TRY(cudaMallocFromPoolAsync((void**)&d_out, es * N * slots * piece, pool, st));
    for (int64_t c0 = lo; c0 < hi; c0 += piece) {
      const int64_t n = std::min<int64_t>(piece, hi - c0);
      SDE_TRY(cudaMemcpy2DAsync(d_u0, es * piece, u0 + es * c0, es * n_all, es * n, N, cudaMemcpyHostToDevice, st));
      if (NP) SDE_TRY(cudaMemcpy2DAsync(d_p, es * piece, p + es * c0, es * n_all, es * n, NP, cudaMemcpyHostToDevice, st));
      if (d_z) SDE_TRY(cudaMemcpy2DAsync(d_z, es * piece, noise + es * c0, es * n_all, es * n, (size_t)n_normals, cudaMemcpyHostToDevice, st));
      sde_em_options_t oc = *o;
      oc.n_traj = n;
      oc.traj_offset = o->traj_offset + c0;     // the Philox counter is the GLOBAL trajectory index
      rc = em_launch(sys, &oc, fn, d_u0, d_p, piece, d_z, piece, d_out, piece, st);
      if (rc != SDE_OK) { cleanup(); return rc; }
      if (o->save_mode == SDE_SAVE_ENDPOINT) {
        SDE_TRY(cudaMemcpy2DAsync(out_u + es * c0, es * n_all, d_out, es * piece, es * n, N, cudaMemcpyDeviceToHost, st));
      } else if (o->layout == SDE_LAYOUT_TRAJ_MAJOR) {
        SDE_TRY(cudaMemcpyAsync(out_u + es * N * slots * c0, d_out, es * N * slots * n, cudaMemcpyDeviceToHost, st));
      } else {
        SDE_TRY(cudaMemcpy2DAsync(out_u + es * c0, es * n_all, d_out, es * piece, es * n, (size_t)N * slots, cudaMemcpyDeviceToHost, st));
      }
      SDE_TRY(cudaStreamSynchronize(st));
    }
#undef SDE_TRY
    cleanup();
    return SDE_OK;
  };
  int rc = body();
  if (rc != SDE_OK && err) *err = g_err;
  return rc;
}

}  // namespace

extern "C" {

int sde_em_system_builtin(const char* name, sde_em_system_t* out) {
  if (!name || !out) return fail(SDE_ERR_INVALID, "null argument");
  std::call_once(g_em_once, init_em_builtins);
  for (int i = 0; i < kNumEmBuiltins; ++i)
    if (!strcmp(name, kEmBuiltins[i].name)) { *out = &g_em_handles[i]; return SDE_OK; }
  return fail(SDE_ERR_INVALID, "unknown built-in SDE system '%s'", name);
}

int sde_em_system_nvrtc(const char* src, int n_state, int n_param, int n_noise, int diagonal, sde_em_system_t* out,
                        char* log, size_t log_len) {
  if (!src || !out) return fail(SDE_ERR_INVALID, "null argument");
  if (n_state < 1 || n_state > 32 || n_param < 0 || n_param > 64 || n_noise < 1 || n_noise > 32)
    return fail(SDE_ERR_INVALID, "n_state and n_noise must be in 1..32, n_param in 0..64");
  if (diagonal && n_noise != n_state) return fail(SDE_ERR_INVALID, "diagonal noise needs n_noise == n_state");
  if (log && log_len) log[0] = '\0';
  sde_em_system_s* s = new sde_em_system_s;
  s->builtin = false;
  s->name = "user";
  s->n_state = n_state; s->n_param = n_param; s->n_noise = n_noise; s->diagonal = diagonal != 0;
  s->src = src;
  for (int dtype = 0; dtype < 2; ++dtype) {
    std::string lg;
    int rc = nvrtc_compile(em_user_program(s, dtype, 0, 0, true), nullptr, &lg);
    if (rc != SDE_OK) {
      if (log && log_len) { strncpy(log, lg.c_str(), log_len - 1); log[log_len - 1] = '\0'; }
      delete s;
      return rc;
    }
  }
  *out = s;
  return SDE_OK;
}

int sde_em_system_dims(sde_em_system_t sys, int* n_state, int* n_param, int* n_noise, int* diagonal) {
  if (!sys) return fail(SDE_ERR_INVALID, "null system");
  if (n_state) *n_state = sys->n_state;
  if (n_param) *n_param = sys->n_param;
  if (n_noise) *n_noise = sys->n_noise;
  if (diagonal) *diagonal = sys->diagonal ? 1 : 0;
  return SDE_OK;
}

void sde_em_system_free(sde_em_system_t sys) {
  if (!sys || sys->builtin) return;
  for (auto& kv : sys->cache)
    if (kv.second.lib) cudaLibraryUnload(kv.second.lib);
  delete sys;
}

int sde_em_system_prepare(sde_em_system_t sys, const sde_em_options_t* opt) {
  int rc = em_validate(sys, opt);
  if (rc != SDE_OK) return rc;
  const void* fn = nullptr;
  return em_get_kernel(sys, opt, false, &fn);
}

int sde_em_solve_device(sde_em_system_t sys, const sde_em_options_t* opt, const void* d_u0, const void* d_p, int64_t ld_in,
                        const void* d_noise, int64_t noise_ld, void* d_out_u, int64_t ld_out, void* stream, int async) {
  int rc = em_validate(sys, opt);
  if (rc != SDE_OK) return rc;
  if (opt->n_traj == 0) return SDE_OK;
  if (!d_u0 || !d_out_u || (sys->n_param > 0 && !d_p)) return fail(SDE_ERR_INVALID, "null device buffer");
  if (ld_in < opt->n_traj || ld_out < opt->n_traj) return fail(SDE_ERR_INVALID, "ld_in / ld_out smaller than n_traj");
  if (opt->noise_mode == SDE_NOISE_PROVIDED && noise_ld < opt->n_traj) return fail(SDE_ERR_INVALID, "noise_ld smaller than n_traj");
  const void* fn = nullptr;
  rc = em_get_kernel(sys, opt, true, &fn);
  if (rc != SDE_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  rc = em_launch(sys, opt, fn, d_u0, d_p, ld_in, d_noise, noise_ld, d_out_u, ld_out, st);
  if (rc != SDE_OK) return rc;
  if (!async) SDE_CUDA(cudaStreamSynchronize(st));
  return SDE_OK;
}

int sde_em_solve(sde_em_system_t sys, const sde_em_options_t* opt, const void* u0, const void* p, const void* noise,
                 void* out_u, const int* devices, int n_dev) {
  int rc = em_validate(sys, opt);
  if (rc != SDE_OK) return rc;
  if (n_dev < 0 || (n_dev > 0 && !devices)) return fail(SDE_ERR_INVALID, "bad device list");
  if (opt->n_traj == 0) return SDE_OK;
  if (!u0 || !out_u || (sys->n_param > 0 && !p)) return fail(SDE_ERR_INVALID, "null host buffer");
  if (opt->noise_mode == SDE_NOISE_PROVIDED && !noise && opt->n_steps > 0)
    return fail(SDE_ERR_INVALID, "noise_mode SDE_NOISE_PROVIDED needs a noise array");
  if (n_dev <= 1)
    return em_solve_shard(sys, opt, n_dev == 1 ? devices[0] : -1, 0, opt->n_traj, (const char*)u0, (const char*)p,
                          (const char*)noise, (char*)out_u, nullptr);
  // contiguous index ranges, one host thread + stream per device, no collective; every step costs the
  // same for every trajectory, so equal ranges are equal work
  std::vector<std::thread> th;
  std::vector<int> rcs(n_dev, SDE_OK);
  std::vector<std::string> errs(n_dev);
  for (int g = 0; g < n_dev; ++g) {
    const int64_t lo = opt->n_traj * g / n_dev, hi = opt->n_traj * (g + 1) / n_dev;
    th.emplace_back([=, &rcs, &errs]() {
      rcs[g] = em_solve_shard(sys, opt, devices[g], lo, hi, (const char*)u0, (const char*)p, (const char*)noise,
                              (char*)out_u, &errs[g]);
    });
  }
  for (auto& t : th) t.join();
  for (int g = 0; g < n_dev; ++g)
    if (rcs[g] != SDE_OK) { g_err = errs[g]; return rcs[g]; }
  return SDE_OK;
}

int sde_em_noise_device(const sde_em_options_t* opt, int n_noise, void* d_out, int64_t ld, void* stream, int async) {
  if (!opt || !d_out) return fail(SDE_ERR_INVALID, "null argument");
  if (opt->dtype != SDE_F64 && opt->dtype != SDE_F32) return fail(SDE_ERR_INVALID, "unknown dtype %d", opt->dtype);
  if (n_noise < 1 || opt->n_traj < 0 || opt->n_steps < 0 || ld < opt->n_traj) return fail(SDE_ERR_INVALID, "bad noise shape");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = em_noise_launch(opt->dtype, opt->seed, opt->traj_offset, opt->n_traj, opt->n_steps * n_noise, d_out, ld, st);
  if (rc != SDE_OK) return rc;
  if (!async) SDE_CUDA(cudaStreamSynchronize(st));
  return SDE_OK;
}

int sde_em_noise(const sde_em_options_t* opt, int n_noise, void* out) {
  if (!opt || !out) return fail(SDE_ERR_INVALID, "null argument");
  if (opt->dtype != SDE_F64 && opt->dtype != SDE_F32) return fail(SDE_ERR_INVALID, "unknown dtype %d", opt->dtype);
  if (n_noise < 1 || opt->n_traj < 0 || opt->n_steps < 0) return fail(SDE_ERR_INVALID, "bad noise shape");
  const size_t bytes = esize(opt->dtype) * (size_t)opt->n_steps * n_noise * (size_t)opt->n_traj;
  if (bytes == 0) return SDE_OK;
  int dev = 0;
  SDE_CUDA(cudaGetDevice(&dev));
  cudaMemPool_t pool;
  int rc = device_pool(dev, &pool);
  if (rc != SDE_OK) return rc;
  cudaStream_t st;
  SDE_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  void* d = nullptr;
  cudaError_t e = cudaMallocFromPoolAsync(&d, bytes, pool, st);
  if (e == cudaSuccess) {
    rc = em_noise_launch(opt->dtype, opt->seed, opt->traj_offset, opt->n_traj, opt->n_steps * n_noise, d, opt->n_traj, st);
    if (rc == SDE_OK) e = cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cudaFreeAsync(d, st);
  }
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  cudaMemPoolTrimTo(pool, pool_keep_bytes());
  if (e != cudaSuccess) return fail(SDE_ERR_CUDA, "sde_em_noise: %s", cudaGetErrorString(e));
  return rc;
}

}  // extern "C"
