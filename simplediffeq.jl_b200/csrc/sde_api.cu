// sde_api.cu -- C ABI of libsimplediffeq_cuda: system registry, NVRTC path, ensemble launcher,
// multi-device sharder.  See include/simplediffeq_cuda.h for the contract of every entry point.
#include <cuda_runtime.h>
#include <nvrtc.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/simplediffeq_cuda.h"
#include "device/sde_common.cuh"
#include "sde_builtin_decl.h"
#include "sde_interp_host_gen.h"

// ids in the public header and in the device headers must agree
static_assert((int)sde::kTsit5 == SDE_ALG_TSIT5 && (int)sde::kATsit5 == SDE_ALG_ATSIT5 &&
              (int)sde::kRK4 == SDE_ALG_RK4 && (int)sde::kVern7 == SDE_ALG_VERN7 &&
              (int)sde::kAVern7 == SDE_ALG_AVERN7 && (int)sde::kVern9 == SDE_ALG_VERN9 &&
              (int)sde::kAVern9 == SDE_ALG_AVERN9 && (int)sde::kEuler == SDE_ALG_EULER, "alg ids");
static_assert((int)sde::kSaveEndpoint == SDE_SAVE_ENDPOINT && (int)sde::kSaveAt == SDE_SAVE_SAVEAT &&
              (int)sde::kSaveEveryStep == SDE_SAVE_EVERYSTEP, "save ids");
static_assert((int)sde::kLayoutTrajMajor == SDE_LAYOUT_TRAJ_MAJOR && (int)sde::kLayoutSoA == SDE_LAYOUT_SOA, "layout ids");
static_assert((int)sde::kRetDefault == SDE_RET_DEFAULT && (int)sde::kRetDtMin == SDE_RET_DTMIN &&
              (int)sde::kRetMaxIters == SDE_RET_MAXITERS && (int)sde::kRetOutputFull == SDE_RET_OUTPUT_FULL, "retcodes");
static_assert((int)sde::kCompatFixVern9Interp == SDE_COMPAT_FIX_VERN9_INTERP &&
              (int)sde::kCompatStrictController == SDE_COMPAT_STRICT_CONTROLLER, "compat flags");

extern const char* const sde_embedded_names[];
extern const char* const sde_embedded_sources[];
extern const int sde_embedded_count;

namespace {

constexpr int kBlock = 128;

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define SDE_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail(SDE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

bool is_adaptive(int alg) { return alg == SDE_ALG_ATSIT5 || alg == SDE_ALG_AVERN7 || alg == SDE_ALG_AVERN9; }

// kernel variant bits (sde_builtin_decl.h)
bool want_q2(const sde_options_t* o) {
  return o->alg == SDE_ALG_VERN9 && o->save_mode == SDE_SAVE_SAVEAT && !(o->compat & SDE_COMPAT_FIX_VERN9_INTERP);
}
bool want_strict(const sde_options_t* o) { return is_adaptive(o->alg) && (o->compat & SDE_COMPAT_STRICT_CONTROLLER); }
bool want_staged(const sde_options_t* o) {
  return !is_adaptive(o->alg) && o->save_mode != SDE_SAVE_ENDPOINT && o->layout == SDE_LAYOUT_TRAJ_MAJOR;
}
int64_t out_slots(const sde_options_t* o) {
  if (o->save_mode == SDE_SAVE_SAVEAT) return o->n_save;
  if (o->save_mode == SDE_SAVE_EVERYSTEP) return is_adaptive(o->alg) ? o->out_capacity : o->n_steps + 1;
  return 1;
}
// dynamic shared memory of the staged writer: must mirror sde::StageCfg
// Tuning knob (NVRTC systems only, development): SDE_TUNE_STAGE_ELEMS=<elements staged per lane>
int tune_stage_elems() {
  const char* e = getenv("SDE_TUNE_STAGE_ELEMS");
  return e ? atoi(e) : 0;
}
size_t staged_smem_bytes(int n_state, size_t es, int block, bool user) {
  int elems = es == 8 ? 45 : 93;
  if (user && tune_stage_elems() > 0) elems = tune_stage_elems();
  const int S = std::max(1, elems / n_state);
  const int LS = (S * n_state) | 1;
  return (size_t)(block / 32) * 32 * LS * es;
}
size_t esize(int dtype) { return dtype == SDE_F64 ? 8 : 4; }

struct Compiled {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kernel = nullptr;
  std::vector<char> cubin;
};

}  // namespace

struct sde_system_s {
  bool builtin = true;
  std::string name;
  int n_state = 0, n_param = 0;
  sde_builtin_lookup_fn lookup = nullptr;
  // NVRTC systems
  std::string src;
  std::mutex mu;
  std::map<std::string, Compiled> cache;   // key: alg/dtype/save/q2[/device]
};

namespace {

struct BuiltinEntry {
  const char* name;
  int n_state, n_param;
  sde_builtin_lookup_fn fn;
};
const BuiltinEntry kBuiltins[] = {
    {"lorenz", 3, 3, sde_lookup_lorenz},
    {"vanderpol", 2, 1, sde_lookup_vanderpol},
    {"robertson", 3, 3, sde_lookup_robertson},
    {"nbody", 12, 3, sde_lookup_nbody},
    {"lineardecay", 3, 3, sde_lookup_lineardecay},
    {"scalargrowth", 1, 1, sde_lookup_scalargrowth},
    {"nonautonomous", 2, 2, sde_lookup_nonautonomous},
};
sde_system_s g_builtin_handles[sizeof(kBuiltins) / sizeof(kBuiltins[0])];
std::once_flag g_builtin_once;

void init_builtins() {
  for (size_t i = 0; i < sizeof(kBuiltins) / sizeof(kBuiltins[0]); ++i) {
    g_builtin_handles[i].builtin = true;
    g_builtin_handles[i].name = kBuiltins[i].name;
    g_builtin_handles[i].n_state = kBuiltins[i].n_state;
    g_builtin_handles[i].n_param = kBuiltins[i].n_param;
    g_builtin_handles[i].lookup = kBuiltins[i].fn;
  }
}

int validate(const sde_system_s* sys, const sde_options_t* o) {
  if (!sys || !o) return fail(SDE_ERR_INVALID, "null system or options");
  if (o->alg < 0 || o->alg > 7) return fail(SDE_ERR_INVALID, "unknown algorithm id %d", o->alg);
  if (o->dtype != SDE_F64 && o->dtype != SDE_F32) return fail(SDE_ERR_INVALID, "unknown dtype %d", o->dtype);
  if (o->save_mode < 0 || o->save_mode > 2) return fail(SDE_ERR_INVALID, "unknown save_mode %d", o->save_mode);
  if (o->layout != SDE_LAYOUT_TRAJ_MAJOR && o->layout != SDE_LAYOUT_SOA)
    return fail(SDE_ERR_INVALID, "unknown layout %d", o->layout);
  if (o->n_traj < 0) return fail(SDE_ERR_INVALID, "n_traj < 0");
  if (o->save_mode == SDE_SAVE_SAVEAT) {
    if (o->n_save < 0 || (o->n_save > 0 && !o->saveat)) return fail(SDE_ERR_INVALID, "saveat array missing");
    if (o->n_save > 0x7fffffff) return fail(SDE_ERR_INVALID, "n_save too large");
    if (o->alg == SDE_ALG_RK4 || o->alg == SDE_ALG_EULER)
      return fail(SDE_ERR_UNSUPPORTED, "GPUSimpleRK4 / GPUSimpleEuler have no saveat (the reference ignores the keyword and saves every step)");
  }
  if (is_adaptive(o->alg)) {
    if (o->save_mode == SDE_SAVE_EVERYSTEP && o->out_capacity < 1)
      return fail(SDE_ERR_INVALID, "adaptive SDE_SAVE_EVERYSTEP needs out_capacity >= 1 (slots per trajectory)");
  } else {
    if (o->n_steps < 0) return fail(SDE_ERR_INVALID, "n_steps < 0");
  }
  return SDE_OK;
}

// --------------------------------------------------------------------------------------------
// NVRTC
// --------------------------------------------------------------------------------------------
const char* method_name(int alg) {
  switch (alg) {
    case SDE_ALG_TSIT5: case SDE_ALG_ATSIT5: return "sde::Tsit5Method";
    case SDE_ALG_RK4: return "sde::RK4Method";
    case SDE_ALG_EULER: return "sde::EulerMethod";
    case SDE_ALG_VERN7: case SDE_ALG_AVERN7: return "sde::Vern7Method";
    default: return "sde::Vern9Method";
  }
}

std::string user_program(const sde_system_s* sys, int alg, int dtype, int save, bool q2, bool strict, bool staged, bool syntax_only) {
  std::string s;
  s += dtype == SDE_F64 ? "typedef double real;\n" : "typedef float real;\n";
  s += "#include \"sde_kernels.cuh\"\n";
  s += "#line 1 \"user_rhs.cu\"\n";
  s += sys->src;
  s += "\n#line 1 \"sde_user_glue.cu\"\n";
  char buf[1024];
  snprintf(buf, sizeof buf,
           "struct SdeUserSys {\n"
           "  static constexpr int N = %d, NP = %d;\n"
           "  template <class T> __device__ __forceinline__ static void rhs(T* du, const T* u, const T* p, T t) {\n"
           "    ::rhs(du, u, p, t);\n  }\n};\n",
           sys->n_state, sys->n_param);
  s += buf;
  if (syntax_only) {
    s += "extern \"C\" __global__ void sde_user_check(real* du, const real* u, const real* p, real t) {\n"
         "  SdeUserSys::rhs<real>(du, u, p, t);\n}\n";
    return s;
  }
  if (is_adaptive(alg)) {
    snprintf(buf, sizeof buf,
             "extern \"C\" __global__ void __launch_bounds__(%d) sde_user_kernel(const __grid_constant__ sde::KArgs<real> a) {\n"
             "  sde::adaptive_body<SdeUserSys, real, %s<SdeUserSys, real>, %d, %s, %s>(a);\n}\n",
             kBlock, method_name(alg), save, alg == SDE_ALG_AVERN9 ? "true" : "false", strict ? "true" : "false");
  } else {
    snprintf(buf, sizeof buf,
             "extern \"C\" __global__ void __launch_bounds__(%d) sde_user_kernel(const __grid_constant__ sde::KArgs<real> a) {\n"
             "  sde::fixed_body<SdeUserSys, real, %s<SdeUserSys, real>, %d, %s, %s>(a);\n}\n",
             kBlock, method_name(alg), save, q2 ? "true" : "false", staged ? "true" : "false");
  }
  s += buf;
  return s;
}

int nvrtc_compile(const std::string& program, std::vector<char>* cubin, std::string* log) {
  nvrtcProgram prog;
  nvrtcResult r = nvrtcCreateProgram(&prog, program.c_str(), "sde_user.cu", sde_embedded_count,
                                     sde_embedded_sources, sde_embedded_names);
  if (r != NVRTC_SUCCESS) return fail(SDE_ERR_NVRTC, "nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
  char tune[96];
  const int te = tune_stage_elems();
  snprintf(tune, sizeof tune, "-DSDE_STAGE_ELEMS_F64=%d", te > 0 ? te : 45);
  char tune32[96];
  snprintf(tune32, sizeof tune32, "-DSDE_STAGE_ELEMS_F32=%d", te > 0 ? te : 93);
  const char* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "--std=c++17", "-lineinfo",
                        "-default-device", tune, tune32};
  r = nvrtcCompileProgram(prog, 7, opts);
  size_t ls = 0;
  nvrtcGetProgramLogSize(prog, &ls);
  std::string lg(ls, '\0');
  if (ls > 1) nvrtcGetProgramLog(prog, &lg[0]);
  if (log) *log = lg;
  if (r != NVRTC_SUCCESS) {
    nvrtcDestroyProgram(&prog);
    return fail(SDE_ERR_NVRTC, "NVRTC compilation failed: %s\n%s", nvrtcGetErrorString(r), lg.c_str());
  }
  if (cubin) {
    size_t cs = 0;
    r = nvrtcGetCUBINSize(prog, &cs);
    if (r != NVRTC_SUCCESS || cs == 0) {
      nvrtcDestroyProgram(&prog);
      return fail(SDE_ERR_NVRTC, "nvrtcGetCUBINSize: %s", nvrtcGetErrorString(r));
    }
    cubin->resize(cs);
    nvrtcGetCUBIN(prog, cubin->data());
  }
  nvrtcDestroyProgram(&prog);
  return SDE_OK;
}

// returns the cache entry (compiled; module loaded only if `load`)
int get_user_kernel(sde_system_s* sys, const sde_options_t* o, bool load, const void** fn) {
  const bool q2 = want_q2(o), strict = want_strict(o), staged = want_staged(o);
  int dev = -1;
  if (load) SDE_CUDA(cudaGetDevice(&dev));
  char key[96];
  snprintf(key, sizeof key, "%d/%d/%d/%d/%d/%d", o->alg, o->dtype, o->save_mode, (int)q2, (int)strict, (int)staged);
  std::lock_guard<std::mutex> lk(sys->mu);
  Compiled& c = sys->cache[key];
  if (c.cubin.empty()) {
    std::string prog = user_program(sys, o->alg, o->dtype, o->save_mode, q2, strict, staged, false);
    int rc = nvrtc_compile(prog, &c.cubin, nullptr);
    if (rc != SDE_OK) { sys->cache.erase(key); return rc; }
  }
  if (!load) return SDE_OK;
  // one module per device
  char dkey[112];
  snprintf(dkey, sizeof dkey, "%s@%d", key, dev);
  Compiled& d = sys->cache[dkey];
  if (!d.kernel) {
    const std::vector<char>& cubin = sys->cache[key].cubin;
    SDE_CUDA(cudaLibraryLoadData(&d.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    SDE_CUDA(cudaLibraryGetKernel(&d.kernel, d.lib, "sde_user_kernel"));
  }
  *fn = (const void*)d.kernel;
  return SDE_OK;
}

int get_kernel(sde_system_s* sys, const sde_options_t* o, bool load, const void** fn) {
  if (sys->builtin) {
    sde::KernelInfo ki = sys->lookup(o->alg, o->dtype, o->save_mode,
                                     (want_q2(o) ? 1 : 0) | (want_strict(o) ? 2 : 0) | (want_staged(o) ? 4 : 0));
    if (!ki.fn)
      return fail(SDE_ERR_UNSUPPORTED, "no kernel for system %s alg %d dtype %d save_mode %d",
                  sys->name.c_str(), o->alg, o->dtype, o->save_mode);
    *fn = ki.fn;
    return SDE_OK;
  }
  return get_user_kernel(sys, o, load, fn);
}

// --------------------------------------------------------------------------------------------
// fixed step + saveat: the schedule of src/tsit5/gpuatsit5.jl:116-127 (`while cur_t <= length(ts) &&
// ts[cur_t] <= t`, theta = (savet - (t - dt))/dt, b(theta) by @evalpoly) does not depend on the
// trajectory.  Evaluate it once here, in T, with the same IEEE operations the device would use.
// --------------------------------------------------------------------------------------------
template <class T>
void build_save_plan(int alg, const T* tgrid, int64_t n_steps, T t0, T dt, const T* saveat, int64_t n_save,
                     std::vector<int>* step, std::vector<T>* b, int* nb_out) {
  const double* poly = nullptr;
  const int* len = nullptr;
  int nb = 0, deg = 0;
  switch (alg) {
    case SDE_ALG_TSIT5: poly = &sde_host::kTsit5Poly[0][0]; len = sde_host::kTsit5Len; nb = sde_host::kTsit5NB; deg = sde_host::kTsit5Deg; break;
    case SDE_ALG_VERN7: poly = &sde_host::kVern7Poly[0][0]; len = sde_host::kVern7Len; nb = sde_host::kVern7NB; deg = sde_host::kVern7Deg; break;
    default: poly = &sde_host::kVern9Poly[0][0]; len = sde_host::kVern9Len; nb = sde_host::kVern9NB; deg = sde_host::kVern9Deg; break;
  }
  *nb_out = nb;
  step->assign((size_t)n_save, (int)std::min<int64_t>(n_steps + 1, 0x7fffffff));   // "never reached"
  b->assign((size_t)n_save * nb, (T)0);
  int64_t cur = 0;
  if (n_save > 0 && t0 == saveat[0]) { (*step)[0] = 0; cur = 1; }
  for (int64_t s = 1; s <= n_steps && cur < n_save; ++s) {
    volatile T tv = tgrid[s - 1];
    tv = tv + dt;                                   // t = _ts[i-1]; t += dt
    const T t = tv;
    while (cur < n_save && saveat[cur] <= t) {
      volatile T tm = t - dt;
      const T th = (saveat[cur] - tm) / dt;
      for (int j = 0; j < nb; ++j) {
        const double* c = poly + (size_t)j * deg;
        T acc = (T)c[len[j] - 1];
        for (int d = len[j] - 2; d >= 0; --d) acc = std::fma(th, acc, (T)c[d]);
        (*b)[(size_t)cur * nb + j] = acc;
      }
      (*step)[(size_t)cur] = (int)s;
      ++cur;
    }
  }
}

// --------------------------------------------------------------------------------------------
// launch on the current device
// --------------------------------------------------------------------------------------------
template <class T>
int launch_t(sde_system_s* sys, const sde_options_t* o, const void* fn, const void* d_u0, const void* d_p,
             int64_t ld_in, void* d_out_u, int64_t ld_out, void* d_out_t, int32_t* d_nacc,
             int32_t* d_nrej, int32_t* d_ret, cudaStream_t st) {
  const bool adaptive = is_adaptive(o->alg);
  sde::KArgs<T> a;
  memset(&a, 0, sizeof a);
  a.u0 = (const T*)d_u0;
  a.p = (const T*)d_p;
  a.n_traj = o->n_traj;
  a.ld_in = ld_in;
  a.t0 = (T)o->t0; a.tf = (T)o->tf; a.dt = (T)o->dt; a.abstol = (T)o->abstol; a.reltol = (T)o->reltol;
  a.n_steps = adaptive ? 0 : o->n_steps;
  a.n_save = o->save_mode == SDE_SAVE_SAVEAT ? (int)o->n_save : 0;
  a.compat = o->compat;
  a.layout = o->layout;
  a.max_attempts = o->max_attempts;
  a.out_u = (T*)d_out_u;
  a.ld_out = ld_out;
  a.n_out = out_slots(o);
  a.out_t = adaptive ? (T*)d_out_t : nullptr;
  a.naccept = d_nacc; a.nreject = d_nrej; a.retcode = d_ret;

  // small per-call device constants: queue head, time grid, saveat (adaptive) or save plan (fixed)
  const size_t ng = adaptive ? 0 : (size_t)o->n_steps + 1;
  const size_t ns = (size_t)a.n_save;
  std::vector<T> tg(ng);
  if (ng) {
    if (o->tgrid) memcpy(tg.data(), o->tgrid, ng * sizeof(T));
    else for (size_t k = 0; k < ng; ++k) tg[k] = (T)o->t0 + (T)((T)k * (T)o->dt);
  }
  std::vector<int> plan_step;
  std::vector<T> plan_b;
  int nb = 0;
  if (!adaptive && ns) {
    if (o->n_steps >= 0x7ffffffeLL) return fail(SDE_ERR_INVALID, "n_steps too large for saveat");
    build_save_plan<T>(o->alg, tg.data(), o->n_steps, (T)o->t0, (T)o->dt, (const T*)o->saveat, o->n_save,
                       &plan_step, &plan_b, &nb);
  }
  auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t off_grid = 16;
  const size_t off_save = off_grid + up16(ng * sizeof(T));
  const size_t off_pstep = off_save + up16((adaptive ? ns : 0) * sizeof(T));
  const size_t off_pb = off_pstep + up16(plan_step.size() * sizeof(int));
  const size_t bytes = off_pb + up16(plan_b.size() * sizeof(T));
  char* scratch = nullptr;
  SDE_CUDA(cudaMallocAsync((void**)&scratch, bytes, st));
  SDE_CUDA(cudaMemsetAsync(scratch, 0, 16, st));
  a.queue = (sde::u64*)scratch;
  // pageable sources are staged by the runtime before cudaMemcpyAsync returns, so the vectors may die
  if (ng) {
    SDE_CUDA(cudaMemcpyAsync(scratch + off_grid, tg.data(), ng * sizeof(T), cudaMemcpyHostToDevice, st));
    a.tgrid = (const T*)(scratch + off_grid);
  }
  if (adaptive && ns) {
    SDE_CUDA(cudaMemcpyAsync(scratch + off_save, o->saveat, ns * sizeof(T), cudaMemcpyHostToDevice, st));
    a.saveat = (const T*)(scratch + off_save);
  }
  if (!plan_step.empty()) {
    SDE_CUDA(cudaMemcpyAsync(scratch + off_pstep, plan_step.data(), plan_step.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    SDE_CUDA(cudaMemcpyAsync(scratch + off_pb, plan_b.data(), plan_b.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    a.plan_step = (const int*)(scratch + off_pstep);
    a.plan_b = (const T*)(scratch + off_pb);
  }

  if (o->n_traj > 0) {
    int dev = 0, sms = 0, per_sm = 0;
    SDE_CUDA(cudaGetDevice(&dev));
    unsigned grid;
    const int64_t full = (o->n_traj + kBlock - 1) / kBlock;
    if (adaptive) {
      SDE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      SDE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kBlock, 0));
      if (per_sm < 1) per_sm = 1;
      grid = (unsigned)std::min<int64_t>(full, (int64_t)sms * per_sm);   // persistent CTAs
    } else {
      if (full > 0x7fffffffLL) return fail(SDE_ERR_INVALID, "n_traj too large for one launch");
      grid = (unsigned)full;
    }
    size_t smem = 0;
    if (want_staged(o)) {
      smem = staged_smem_bytes(sys->n_state, sizeof(T), kBlock, !sys->builtin);
      if (smem > 48 * 1024)
        SDE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    void* params[] = {&a};
    SDE_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(kBlock), params, smem, st));
    g_launches.fetch_add(1);
  }
  SDE_CUDA(cudaFreeAsync(scratch, st));
  return SDE_OK;
}

int launch(sde_system_s* sys, const sde_options_t* o, const void* d_u0, const void* d_p, int64_t ld_in,
           void* d_out_u, int64_t ld_out, void* d_out_t, int32_t* d_nacc, int32_t* d_nrej,
           int32_t* d_ret, cudaStream_t st) {
  const void* fn = nullptr;
  int rc = get_kernel(sys, o, true, &fn);
  if (rc != SDE_OK) return rc;
  if (o->dtype == SDE_F64)
    return launch_t<double>(sys, o, fn, d_u0, d_p, ld_in, d_out_u, ld_out, d_out_t, d_nacc, d_nrej, d_ret, st);
  return launch_t<float>(sys, o, fn, d_u0, d_p, ld_in, d_out_u, ld_out, d_out_t, d_nacc, d_nrej, d_ret, st);
}


// --------------------------------------------------------------------------------------------
// Source of contiguous trajectory ranges for one device of a host-buffer solve.
//   static : the device owns [lo, hi) = [floor(gN/G), floor((g+1)N/G)) and walks it in pieces that
//            fit its memory budget (fixed-step algorithms: every trajectory costs the same);
//   dynamic: all devices pull pieces of `grain` trajectories from one shared counter (adaptive
//            algorithms: step counts vary 10x along a parameter sweep, so equal index ranges are
//            not equal work -- the host-level analogue of the kernels' work queue).
// --------------------------------------------------------------------------------------------
struct RangeSource {
  int64_t lo = 0, hi = 0;                    // static range (cursor = lo)
  std::atomic<int64_t>* shared = nullptr;    // dynamic: next unassigned trajectory
  int64_t total = 0, grain = 0;
  bool next(int64_t max_len, int64_t* a, int64_t* b) {
    if (shared) {
      const int64_t len = std::min(max_len, grain);
      const int64_t s = shared->fetch_add(len);
      if (s >= total) return false;
      *a = s; *b = std::min(total, s + len);
      return true;
    }
    if (lo >= hi) return false;
    *a = lo; *b = std::min(hi, lo + max_len);
    lo = *b;
    return true;
  }
  int64_t max_piece() const { return shared ? std::min(grain, total) : hi - lo; }
};

// one device of a host-buffer solve
int solve_shard(sde_system_s* sys, const sde_options_t* o, int device, RangeSource src,
                const char* u0, const char* p, char* out_u, char* out_t, int32_t* nacc, int32_t* nrej,
                int32_t* ret, std::string* err) {
  auto body = [&]() -> int {
    if (device >= 0) SDE_CUDA(cudaSetDevice(device));
    const size_t es = esize(o->dtype);
    const int N = sys->n_state, NP = sys->n_param;
    const int64_t n_all = o->n_traj, slots = out_slots(o);
    const bool series = o->save_mode != SDE_SAVE_ENDPOINT;
    const bool adaptive = is_adaptive(o->alg);
    // device memory budget per chunk: keep well below HBM capacity
    size_t free_b = 0, total_b = 0;
    SDE_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t per_traj = es * ((size_t)N + NP + (size_t)(N + 1) * slots + 1) + 12;
    int64_t chunk = (int64_t)std::max<size_t>(1, (size_t)(0.6 * (double)free_b) / per_traj);
    chunk = std::min<int64_t>(chunk, src.max_piece());
    if (chunk > 32) chunk -= chunk % 32;
    cudaStream_t st;
    SDE_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    char *d_u0 = nullptr, *d_p = nullptr, *d_out = nullptr, *d_t = nullptr;
    int32_t *d_na = nullptr, *d_nr = nullptr, *d_rc = nullptr;
    int rc = SDE_OK;
    auto cleanup = [&]() {
      cudaFree(d_u0); cudaFree(d_p); cudaFree(d_out); cudaFree(d_t); cudaFree(d_na); cudaFree(d_nr); cudaFree(d_rc);
      cudaStreamDestroy(st);
    };
#define SDE_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); \
      return fail(SDE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
    if (chunk > 0) {
      SDE_TRY(cudaMalloc((void**)&d_u0, es * N * chunk));
      if (NP) SDE_TRY(cudaMalloc((void**)&d_p, es * NP * chunk));
      SDE_TRY(cudaMalloc((void**)&d_out, es * N * slots * chunk));
      const bool t_series = adaptive && o->save_mode == SDE_SAVE_EVERYSTEP;
      if (adaptive && out_t) SDE_TRY(cudaMalloc((void**)&d_t, es * chunk * (t_series ? slots : 1)));
      if (nacc) SDE_TRY(cudaMalloc((void**)&d_na, 4 * chunk));
      if (nrej) SDE_TRY(cudaMalloc((void**)&d_nr, 4 * chunk));
      if (ret) SDE_TRY(cudaMalloc((void**)&d_rc, 4 * chunk));
    }
    int64_t c0 = 0, c1 = 0;
    while (chunk > 0 && src.next(chunk, &c0, &c1)) {
      const int64_t n = c1 - c0;
      // H2D: SoA rows of the shard (host pitch = n_all elements, device pitch = chunk elements)
      SDE_TRY(cudaMemcpy2DAsync(d_u0, es * chunk, u0 + es * c0, es * n_all, es * n, N, cudaMemcpyHostToDevice, st));
      if (NP) SDE_TRY(cudaMemcpy2DAsync(d_p, es * chunk, p + es * c0, es * n_all, es * n, NP, cudaMemcpyHostToDevice, st));
      sde_options_t oc = *o;
      oc.n_traj = n;
      rc = launch(sys, &oc, d_u0, d_p, chunk, d_out, chunk, d_t, d_na, d_nr, d_rc, st);
      if (rc != SDE_OK) { cleanup(); return rc; }
      // D2H
      if (!series) {
        SDE_TRY(cudaMemcpy2DAsync(out_u + es * c0, es * n_all, d_out, es * chunk, es * n, N, cudaMemcpyDeviceToHost, st));
      } else if (o->layout == SDE_LAYOUT_TRAJ_MAJOR) {
        SDE_TRY(cudaMemcpyAsync(out_u + es * N * slots * c0, d_out, es * N * slots * n, cudaMemcpyDeviceToHost, st));
      } else {
        SDE_TRY(cudaMemcpy2DAsync(out_u + es * c0, es * n_all, d_out, es * chunk, es * n, (size_t)N * slots, cudaMemcpyDeviceToHost, st));
      }
      if (d_t) {
        const bool t_series = adaptive && o->save_mode == SDE_SAVE_EVERYSTEP;
        if (!t_series) SDE_TRY(cudaMemcpyAsync(out_t + es * c0, d_t, es * n, cudaMemcpyDeviceToHost, st));
        else if (o->layout == SDE_LAYOUT_TRAJ_MAJOR)
          SDE_TRY(cudaMemcpyAsync(out_t + es * slots * c0, d_t, es * slots * n, cudaMemcpyDeviceToHost, st));
        else
          SDE_TRY(cudaMemcpy2DAsync(out_t + es * c0, es * n_all, d_t, es * chunk, es * n, (size_t)slots, cudaMemcpyDeviceToHost, st));
      }
      if (d_na) SDE_TRY(cudaMemcpyAsync(nacc + c0, d_na, 4 * n, cudaMemcpyDeviceToHost, st));
      if (d_nr) SDE_TRY(cudaMemcpyAsync(nrej + c0, d_nr, 4 * n, cudaMemcpyDeviceToHost, st));
      if (d_rc) SDE_TRY(cudaMemcpyAsync(ret + c0, d_rc, 4 * n, cudaMemcpyDeviceToHost, st));
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

// ==============================================================================================
// exported functions
// ==============================================================================================
extern "C" {

int sde_version(void) { return SDE_VERSION; }

const char* sde_last_error(void) { return g_err.c_str(); }

int sde_device_count(int* count) {
  if (!count) return fail(SDE_ERR_INVALID, "null count");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(SDE_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  *count = n;
  return SDE_OK;
}

int sde_system_builtin(const char* name, sde_system_t* out) {
  if (!name || !out) return fail(SDE_ERR_INVALID, "null argument");
  std::call_once(g_builtin_once, init_builtins);
  for (size_t i = 0; i < sizeof(kBuiltins) / sizeof(kBuiltins[0]); ++i)
    if (!strcmp(name, kBuiltins[i].name)) {
      *out = &g_builtin_handles[i];
      return SDE_OK;
    }
  return fail(SDE_ERR_INVALID, "unknown built-in system '%s'", name);
}

int sde_system_nvrtc(const char* src, int n_state, int n_param, sde_system_t* out, char* log, size_t log_len) {
  if (!src || !out) return fail(SDE_ERR_INVALID, "null argument");
  if (n_state < 1 || n_state > 64 || n_param < 0 || n_param > 64)
    return fail(SDE_ERR_INVALID, "n_state must be in 1..64 and n_param in 0..64");
  if (log && log_len) log[0] = '\0';
  sde_system_s* s = new sde_system_s;
  s->builtin = false;
  s->name = "user";
  s->n_state = n_state;
  s->n_param = n_param;
  s->src = src;
  // syntax check now (both element types must compile), kernels are built lazily
  for (int dtype = 0; dtype < 2; ++dtype) {
    std::string lg;
    int rc = nvrtc_compile(user_program(s, 0, dtype, 0, false, false, false, true), nullptr, &lg);
    if (rc != SDE_OK) {
      if (log && log_len) { strncpy(log, lg.c_str(), log_len - 1); log[log_len - 1] = '\0'; }
      delete s;
      return rc;
    }
  }
  *out = s;
  return SDE_OK;
}

int sde_system_dims(sde_system_t sys, int* n_state, int* n_param) {
  if (!sys) return fail(SDE_ERR_INVALID, "null system");
  if (n_state) *n_state = sys->n_state;
  if (n_param) *n_param = sys->n_param;
  return SDE_OK;
}

void sde_system_free(sde_system_t sys) {
  if (!sys || sys->builtin) return;
  for (auto& kv : sys->cache)
    if (kv.second.lib) cudaLibraryUnload(kv.second.lib);
  delete sys;
}

int sde_system_prepare(sde_system_t sys, const sde_options_t* opt) {
  int rc = validate(sys, opt);
  if (rc != SDE_OK) return rc;
  const void* fn = nullptr;
  return get_kernel(sys, opt, false, &fn);
}

int sde_solve_device(sde_system_t sys, const sde_options_t* opt, const void* d_u0, const void* d_p,
                     int64_t ld_in, void* d_out_u, int64_t ld_out, void* d_out_t, int32_t* d_naccept,
                     int32_t* d_nreject, int32_t* d_retcode, void* stream, int async) {
  int rc = validate(sys, opt);
  if (rc != SDE_OK) return rc;
  if (!d_u0 || !d_out_u || (sys->n_param > 0 && !d_p)) return fail(SDE_ERR_INVALID, "null device buffer");
  if (ld_in < opt->n_traj || ld_out < opt->n_traj) return fail(SDE_ERR_INVALID, "ld_in / ld_out smaller than n_traj");
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch(sys, opt, d_u0, d_p, ld_in, d_out_u, ld_out, d_out_t, d_naccept, d_nreject, d_retcode, st);
  if (rc != SDE_OK) return rc;
  if (!async) SDE_CUDA(cudaStreamSynchronize(st));
  return SDE_OK;
}

int sde_solve(sde_system_t sys, const sde_options_t* opt, const void* u0, const void* p, void* out_u,
              void* out_t, int32_t* naccept, int32_t* nreject, int32_t* retcode, const int* devices,
              int n_dev) {
  int rc = validate(sys, opt);
  if (rc != SDE_OK) return rc;
  if (!u0 || !out_u || (sys->n_param > 0 && !p)) return fail(SDE_ERR_INVALID, "null host buffer");
  if (n_dev < 0 || (n_dev > 0 && !devices)) return fail(SDE_ERR_INVALID, "bad device list");
  if (opt->n_traj == 0) return SDE_OK;
  if (n_dev <= 1) {
    RangeSource all;
    all.lo = 0; all.hi = opt->n_traj;
    return solve_shard(sys, opt, n_dev == 1 ? devices[0] : -1, all, (const char*)u0,
                       (const char*)p, (char*)out_u, (char*)out_t, naccept, nreject, retcode, nullptr);
  }
  // contiguous index ranges [g*N/G, (g+1)*N/G), one host thread + stream per device, no collective
  // (adaptive algorithms: pieces of `grain` trajectories from a shared counter instead, see RangeSource)
  std::vector<std::thread> th;
  std::vector<int> rcs(n_dev, SDE_OK);
  std::vector<std::string> errs(n_dev);
  std::atomic<int64_t> cursor{0};
  for (int g = 0; g < n_dev; ++g) {
    RangeSource src;
    if (is_adaptive(opt->alg) && !getenv("SDE_TUNE_STATIC_SHARDS")) {   // env knob: measurement only
      src.shared = &cursor;
      src.total = opt->n_traj;
      int64_t grain = opt->n_traj / ((int64_t)n_dev * 8);          // ~8 pieces per device
      grain = std::max<int64_t>(grain, 1 << 16);                     // but never tiny launches
      src.grain = (grain + 31) & ~(int64_t)31;
    } else {
      src.lo = opt->n_traj * g / n_dev;
      src.hi = opt->n_traj * (g + 1) / n_dev;
    }
    th.emplace_back([=, &rcs, &errs]() {
      rcs[g] = solve_shard(sys, opt, devices[g], src, (const char*)u0, (const char*)p, (char*)out_u,
                           (char*)out_t, naccept, nreject, retcode, &errs[g]);
    });
  }
  for (auto& t : th) t.join();
  for (int g = 0; g < n_dev; ++g)
    if (rcs[g] != SDE_OK) return fail(rcs[g], "device %d: %s", devices[g], errs[g].c_str());
  return SDE_OK;
}

int sde_fixed_times(const sde_options_t* o, void* out, int64_t n, int64_t* n_written) {
  if (!o || !out || !n_written) return fail(SDE_ERR_INVALID, "null argument");
  if (is_adaptive(o->alg)) return fail(SDE_ERR_INVALID, "sde_fixed_times is for fixed-step algorithms");
  const int64_t need = o->save_mode == SDE_SAVE_SAVEAT ? o->n_save
                       : o->save_mode == SDE_SAVE_EVERYSTEP ? o->n_steps + 1 : 2;
  if (n < need) return fail(SDE_ERR_INVALID, "output too small: need %lld", (long long)need);
  auto run = [&](auto zero) {
    using T = decltype(zero);
    T* t = (T*)out;
    const T* g = (const T*)o->tgrid;
    auto grid = [&](int64_t k) -> T { return g ? g[k] : (T)((T)o->t0 + (T)((T)k * (T)o->dt)); };
    const T dt = (T)o->dt;
    if (o->save_mode == SDE_SAVE_SAVEAT) {
      memcpy(t, o->saveat, sizeof(T) * o->n_save);
    } else if (o->alg == SDE_ALG_RK4 || o->alg == SDE_ALG_EULER) {
      // ts = tspan[1]:dt:tspan[2] itself (src/rk4/gpurk4.jl:65, src/euler/gpueuler.jl:66)
      if (o->save_mode == SDE_SAVE_EVERYSTEP) for (int64_t k = 0; k <= o->n_steps; ++k) t[k] = grid(k);
      else { t[0] = (T)o->t0; t[1] = grid(o->n_steps); }
    } else {
      // t = _ts[i-1]; ...; t += dt; push!(ts, t)   (src/tsit5/gpuatsit5.jl:98,111,114)
      if (o->save_mode == SDE_SAVE_EVERYSTEP) {
        t[0] = (T)o->t0;
        for (int64_t k = 1; k <= o->n_steps; ++k) t[k] = (T)(grid(k - 1) + dt);
      } else {
        t[0] = (T)o->t0;
        t[1] = o->n_steps > 0 ? (T)(grid(o->n_steps - 1) + dt) : (T)o->t0;
      }
    }
  };
  if (o->dtype == SDE_F64) run(0.0); else run(0.0f);
  *n_written = need;
  return SDE_OK;
}

int sde_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return fail(SDE_ERR_INVALID, "null ptr");
  SDE_CUDA(cudaMallocHost(ptr, bytes));
  return SDE_OK;
}

int sde_host_free(void* ptr) {
  SDE_CUDA(cudaFreeHost(ptr));
  return SDE_OK;
}

int64_t sde_launch_count(void) { return g_launches.load(); }

}  // extern "C"
