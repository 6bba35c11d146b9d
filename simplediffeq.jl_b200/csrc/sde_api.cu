// sde_api.cu -- C ABI of libsimplediffeq_cuda: system registry, NVRTC path, ensemble launcher,
// multi-device sharder.  See include/simplediffeq_cuda.h for the contract of every entry point.
#include <cuda_runtime.h>
#include <nvrtc.h>
#include <chrono>
#include <sys/stat.h>
#include <unistd.h>

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
#include "sde_internal.h"
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
              (int)sde::kCompatStrictController == SDE_COMPAT_STRICT_CONTROLLER &&
              (int)sde::kCompatLog2Controller == SDE_COMPAT_LOG2_CONTROLLER &&
              (int)sde::kCompatFastRhs == SDE_COMPAT_FAST_RHS && (int)sde::kCompatFastStages == SDE_COMPAT_FAST_STAGES, "compat flags");

extern const char* const sde_embedded_names[];
extern const char* const sde_embedded_sources[];
extern const int sde_embedded_count;

namespace sde_host {

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

// dynamic shared memory of the staged writer: must mirror sde::StageCfg
// Tuning knob (NVRTC systems only, development): SDE_TUNE_STAGE_ELEMS=<elements staged per lane>
int tune_stage_elems() {
  const char* e = getenv("SDE_TUNE_STAGE_ELEMS");
  return e ? atoi(e) : 0;
}
// SDE_TUNE_MIN_BLOCKS=<n> (development, NVRTC systems only): second __launch_bounds__ argument of the fixed-step kernels
int tune_min_blocks() {
  const char* e = getenv("SDE_TUNE_MIN_BLOCKS");
  return e ? atoi(e) : 0;
}
size_t staged_smem_bytes(int n_state, size_t es, int block, bool user) {
  int elems = es == 8 ? 48 : 96;
  if (user && tune_stage_elems() > 0) elems = tune_stage_elems();
  // StageCfg::kCapB: the configured capacity, at least one line + line offset + one slot, an odd number of 16-byte units
  const int sz = (int)es;
  const int need = (128 - sz) + 128 + (n_state * sz - sz);
  const int want = elems * sz + 16;
  const int raw = (std::max(want, need) + 15) / 16 * 16;
  const int cap = ((raw / 16) % 2 == 0) ? raw + 16 : raw;
  // per warp: 32 lane regions + the 1 KB ring of dense-output weights (StageCfg::kBytesPerWarp)
  return (size_t)(block / 32) * ((size_t)32 * cap + 1024);
}

}  // namespace sde_host
using namespace sde_host;

namespace {


bool is_adaptive(int alg) { return alg == SDE_ALG_ATSIT5 || alg == SDE_ALG_AVERN7 || alg == SDE_ALG_AVERN9; }

// kernel variant bits (sde_builtin_decl.h)
bool want_q2(const sde_options_t* o) {
  return o->alg == SDE_ALG_VERN9 && o->save_mode == SDE_SAVE_SAVEAT && !(o->compat & SDE_COMPAT_FIX_VERN9_INTERP);
}
// literal or log2-domain step-size controller (see the compat flags in simplediffeq_cuda.h): forced by a flag,
// otherwise literal exactly where the log2 one cannot reproduce the reference's step counts
bool want_strict(const sde_options_t* o) {
  if (!is_adaptive(o->alg)) return false;
  if (o->compat & SDE_COMPAT_STRICT_CONTROLLER) return true;
  if (o->compat & SDE_COMPAT_LOG2_CONTROLLER) return false;
  return o->dtype == SDE_F32 || o->reltol <= 1e-11;
}
// SDE_COMPAT_FAST_STAGES acts on the fixed-step Tsit5 kernels (Tsit5FastMethod), whatever they save
bool want_fast_stages(const sde_options_t* o) {
  return (o->compat & SDE_COMPAT_FAST_STAGES) && o->alg == SDE_ALG_TSIT5;
}
bool want_staged(const sde_options_t* o) {
  return !is_adaptive(o->alg) && o->save_mode != SDE_SAVE_ENDPOINT && o->layout == SDE_LAYOUT_TRAJ_MAJOR;
}
// adaptive save_everystep in the trajectory-major layout: rows through a two-line shared-memory ring per lane
// (sde::adaptive_body; the kernel derives the same condition from its template arguments and KArgs::layout)
bool want_row_stage(const sde_options_t* o, int n_state) {
  return is_adaptive(o->alg) && o->save_mode == SDE_SAVE_EVERYSTEP && o->layout == SDE_LAYOUT_TRAJ_MAJOR &&
         n_state * (o->dtype == SDE_F64 ? 8 : 4) <= sde::kRowStageMaxSlotBytes;
}
int64_t out_slots(const sde_options_t* o) {
  if (o->save_mode == SDE_SAVE_SAVEAT) return o->n_save;
  if (o->save_mode == SDE_SAVE_EVERYSTEP) return is_adaptive(o->alg) ? o->out_capacity : o->n_steps + 1;
  return 1;
}
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
  sde_builtin_lookup_fn lookup_fast = nullptr;   // SDE_COMPAT_FAST_RHS twin (contracted right-hand side), if any
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
  sde_builtin_lookup_fn fn_fast;   // SDE_COMPAT_FAST_RHS twin or null
};
const BuiltinEntry kBuiltins[] = {
    {"lorenz", 3, 3, sde_lookup_lorenz, sde_lookup_lorenz_fma},
    {"vanderpol", 2, 1, sde_lookup_vanderpol, sde_lookup_vanderpol_fma},
    {"robertson", 3, 3, sde_lookup_robertson, nullptr},
    {"nbody", 12, 3, sde_lookup_nbody, nullptr},
    {"lineardecay", 3, 3, sde_lookup_lineardecay, nullptr},
    {"scalargrowth", 1, 1, sde_lookup_scalargrowth, nullptr},
    {"nonautonomous", 2, 2, sde_lookup_nonautonomous, nullptr},
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
    g_builtin_handles[i].lookup_fast = kBuiltins[i].fn_fast;
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
  if (o->compat & ~(SDE_COMPAT_FIX_VERN9_INTERP | SDE_COMPAT_STRICT_CONTROLLER | SDE_COMPAT_LOG2_CONTROLLER | SDE_COMPAT_FAST_RHS | SDE_COMPAT_FAST_STAGES))
    return fail(SDE_ERR_INVALID, "unknown compat flags 0x%x", (unsigned)o->compat);
  if ((o->compat & SDE_COMPAT_STRICT_CONTROLLER) && (o->compat & SDE_COMPAT_LOG2_CONTROLLER))
    return fail(SDE_ERR_INVALID, "compat flags SDE_COMPAT_STRICT_CONTROLLER and SDE_COMPAT_LOG2_CONTROLLER exclude each other");
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

std::string user_program(const sde_system_s* sys, int alg, int dtype, int save, bool q2, bool strict, bool staged, bool syntax_only, bool fast_stages = false) {
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
             "extern \"C\" __global__ void __launch_bounds__(%d, %d) sde_user_kernel(const __grid_constant__ sde::KArgs<real> a) {\n"
             "  sde::fixed_body<SdeUserSys, real, %s<SdeUserSys, real>, %d, %s, %s>(a);\n}\n",
             kBlock, tune_min_blocks() > 0 ? tune_min_blocks() : ((staged && !(((alg == SDE_ALG_VERN7 || alg == SDE_ALG_VERN9) && dtype == SDE_F64) || sys->n_state > 4)) ? 4 : 1),
             // SDE_COMPAT_FAST_STAGES, fixed-step Tsit5: step size folded into the stage coefficients
             (fast_stages && alg == SDE_ALG_TSIT5) ? "sde::Tsit5FastMethod" : method_name(alg), save, q2 ? "true" : "false", staged ? "true" : "false");
  }
  s += buf;
  return s;
}

}  // namespace
namespace sde_host {

// ---- on-disk cubin cache (keyed by a hash of everything that determines the cubin) -----------------
// SDE_CACHE_DIR = directory (unset, empty or "off": no disk cache -- the library never writes outside a
// directory it was given).  Makes a user RHS cost one NVRTC compile per (source, algorithm, dtype, save
// mode) per machine instead of per process.
static std::string cache_dir() {
  const char* e = getenv("SDE_CACHE_DIR");
  return (!e || !*e || !strcmp(e, "off")) ? std::string() : std::string(e);
}
static void fnv(unsigned long long* h, const void* data, size_t n) {
  const unsigned char* p = (const unsigned char*)data;
  for (size_t i = 0; i < n; ++i) { *h ^= p[i]; *h *= 1099511628211ULL; }
}
static std::string cache_key(const std::string& program, const char* const* opts, int n_opts) {
  unsigned long long a = 14695981039346656037ULL, b = 0x9E3779B97F4A7C15ULL;
  auto mix = [&](const void* d, size_t n) { fnv(&a, d, n); fnv(&b, d, n); b = (b << 13) | (b >> 51); };
  int major = 0, minor = 0;
  nvrtcVersion(&major, &minor);
  const int ver[3] = {SDE_VERSION, major, minor};
  mix(ver, sizeof ver);
  mix(program.data(), program.size());
  for (int i = 0; i < n_opts; ++i) mix(opts[i], strlen(opts[i]) + 1);
  for (int i = 0; i < sde_embedded_count; ++i) mix(sde_embedded_sources[i], strlen(sde_embedded_sources[i]) + 1);
  char buf[40];
  snprintf(buf, sizeof buf, "%016llx%016llx", a, b);
  return buf;
}
static bool cache_load(const std::string& path, std::vector<char>* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  bool ok = n > 0;
  if (ok) { out->resize((size_t)n); ok = fread(out->data(), 1, (size_t)n, f) == (size_t)n; }
  fclose(f);
  if (!ok) out->clear();
  return ok;
}
static void cache_store(const std::string& dir, const std::string& path, const std::vector<char>& data) {
  std::string cmd_dir = dir;
  for (size_t i = 1; i <= cmd_dir.size(); ++i)           // mkdir -p
    if (i == cmd_dir.size() || cmd_dir[i] == '/') { std::string sub = cmd_dir.substr(0, i); mkdir(sub.c_str(), 0755); }
  char tmp[64];
  snprintf(tmp, sizeof tmp, ".tmp.%ld.%p", (long)getpid(), (const void*)&data);
  const std::string t = dir + "/" + tmp;
  FILE* f = fopen(t.c_str(), "wb");
  if (!f) return;
  const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
  fclose(f);
  if (!ok || rename(t.c_str(), path.c_str()) != 0) remove(t.c_str());   // atomic publish
}

int nvrtc_compile(const std::string& program, std::vector<char>* cubin, std::string* log, bool fmad) {
  const bool trace = getenv("SDE_TRACE") != nullptr;
  char tune_k[96], tune32_k[96];
  {
    const int te = tune_stage_elems();
    snprintf(tune_k, sizeof tune_k, "-DSDE_STAGE_ELEMS_F64=%d", te > 0 ? te : 48);
    snprintf(tune32_k, sizeof tune32_k, "-DSDE_STAGE_ELEMS_F32=%d", te > 0 ? te : 96);
  }
  // --fmad=false: only explicit fma() fuses (the reference's @muladd placement; the user's f rounds as written).
  // SDE_COMPAT_FAST_RHS compiles the program with --fmad=true instead.
  // SDE_TUNE_DEFINES="-DNAME=value ..." (development, NVRTC systems only): extra macro definitions for A/B runs of
  // kernel variants in one process
  std::vector<std::string> extra;
  if (const char* e = getenv("SDE_TUNE_DEFINES")) {
    std::string cur_opt;
    for (const char* c = e;; ++c) {
      if (*c == ' ' || *c == '\0') { if (cur_opt.rfind("-D", 0) == 0) extra.push_back(cur_opt); cur_opt.clear(); if (!*c) break; }
      else cur_opt += *c;
    }
  }
  std::vector<const char*> key_vec = {"--gpu-architecture=sm_100a", fmad ? "--fmad=true" : "--fmad=false", "--std=c++17", "-lineinfo", "-default-device", tune_k, tune32_k,
                                     "-DSDE_RING_CPASYNC=1"};     // weight ring of the staged kernels: see sde_kernels.cuh
  for (const std::string& x : extra) key_vec.push_back(x.c_str());
  const char* const* key_opts = key_vec.data();
  const int n_key_opts = (int)key_vec.size();
  std::string dir, path;
  if (cubin) {
    dir = cache_dir();
    if (!dir.empty()) {
      path = dir + "/" + cache_key(program, key_opts, n_key_opts) + ".cubin";
      if (cache_load(path, cubin)) {
        if (trace) fprintf(stderr, "[sde trace] nvrtc: cache hit %s\n", path.c_str());
        if (log) log->clear();
        return SDE_OK;
      }
    }
  }
  const auto t_begin = std::chrono::steady_clock::now();
  nvrtcProgram prog;
  nvrtcResult r = nvrtcCreateProgram(&prog, program.c_str(), "sde_user.cu", sde_embedded_count,
                                     sde_embedded_sources, sde_embedded_names);
  if (r != NVRTC_SUCCESS) return fail(SDE_ERR_NVRTC, "nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
  r = nvrtcCompileProgram(prog, n_key_opts, key_opts);
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
    if (!path.empty()) cache_store(dir, path, *cubin);
    if (trace)
      fprintf(stderr, "[sde trace] nvrtc: compiled %zu bytes in %.0f ms%s\n", cs,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(),
              path.empty() ? "" : " (cached)");
  }
  nvrtcDestroyProgram(&prog);
  return SDE_OK;
}
}  // namespace sde_host
namespace {

// returns the cache entry (compiled; module loaded only if `load`)
int get_user_kernel(sde_system_s* sys, const sde_options_t* o, bool load, const void** fn) {
  const bool q2 = want_q2(o), strict = want_strict(o), staged = want_staged(o);
  const bool fast = (o->compat & SDE_COMPAT_FAST_RHS) != 0;
  const bool fast_stages = want_fast_stages(o);
  int dev = -1;
  if (load) SDE_CUDA(cudaGetDevice(&dev));
  char key[96];
  snprintf(key, sizeof key, "%d/%d/%d/%d/%d/%d/%d/%d", o->alg, o->dtype, o->save_mode, (int)q2, (int)strict, (int)staged, (int)fast, (int)fast_stages);
  std::lock_guard<std::mutex> lk(sys->mu);
  Compiled& c = sys->cache[key];
  if (c.cubin.empty()) {
    std::string prog = user_program(sys, o->alg, o->dtype, o->save_mode, q2, strict, staged, false, fast_stages);
    int rc = nvrtc_compile(prog, &c.cubin, nullptr, fast);
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
    const sde_builtin_lookup_fn look = ((o->compat & SDE_COMPAT_FAST_RHS) && sys->lookup_fast) ? sys->lookup_fast : sys->lookup;
    sde::KernelInfo ki = look(o->alg, o->dtype, o->save_mode,
                                     (want_q2(o) ? 1 : 0) | (want_strict(o) ? 2 : 0) | (want_staged(o) ? 4 : 0) | (want_fast_stages(o) ? 8 : 0));
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
  // rows of nbp = nb rounded up to 16 bytes: the kernels read a save point's weights with 16-byte vector loads
  // (sde::plan_stride)
  const int va = (int)(16 / sizeof(T));
  const int nbp = (nb + va - 1) / va * va;
  *nb_out = nbp;
  step->assign((size_t)n_save, (int)std::min<int64_t>(n_steps + 1, 0x7fffffff));   // "never reached"
  b->assign((size_t)n_save * nbp, (T)0);
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
        (*b)[(size_t)cur * nbp + j] = acc;
      }
      (*step)[(size_t)cur] = (int)s;
      ++cur;
    }
  }
}

// --------------------------------------------------------------------------------------------
// device memory: one explicit stream-ordered pool per device.  Freed blocks stay cached in the pool
// (release threshold = max), so repeated solves do not pay cudaMalloc / cudaFree; after a
// host-buffer solve the pool is trimmed to SDE_POOL_KEEP_MB (default 4096 MB), and sde_trim()
// returns everything.
// --------------------------------------------------------------------------------------------
}  // namespace
namespace sde_host {
std::mutex g_pool_mu;
std::map<int, cudaMemPool_t> g_pools;

int device_pool(int dev, cudaMemPool_t* out) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  auto it = g_pools.find(dev);
  if (it != g_pools.end()) { *out = it->second; return SDE_OK; }
  cudaMemPoolProps props;
  memset(&props, 0, sizeof props);
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = dev;
  cudaMemPool_t pool;
  SDE_CUDA(cudaMemPoolCreate(&pool, &props));
  unsigned long long keep_all = ~0ULL;
  SDE_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep_all));
  g_pools[dev] = pool;
  *out = pool;
  return SDE_OK;
}

size_t pool_keep_bytes() {
  const char* e = getenv("SDE_POOL_KEEP_MB");
  const long long mb = e ? atoll(e) : 4096;
  return (size_t)std::max<long long>(0, mb) << 20;
}
}  // namespace sde_host
namespace {

// --------------------------------------------------------------------------------------------
// per-solve device constants (identical for every piece of a chunked solve): time grid, saveat
// (adaptive) or save plan (fixed), plus kQueueSlots work-queue heads (one per in-flight launch)
// --------------------------------------------------------------------------------------------
constexpr int kQueueSlots = 4;

struct SolveConsts {
  char* scratch = nullptr;
  const void* tgrid = nullptr;
  const void* saveat = nullptr;
  const int* plan_cnt = nullptr;
  const void* plan_b = nullptr;
  sde::u64* queue(int slot) const { return (sde::u64*)(scratch + 16 * slot); }
};

template <class T>
int upload_consts_t(const sde_options_t* o, cudaMemPool_t pool, cudaStream_t st, SolveConsts* c) {
  const bool adaptive = is_adaptive(o->alg);
  const size_t ng = adaptive ? 0 : (size_t)o->n_steps + 1;
  const size_t ns = o->save_mode == SDE_SAVE_SAVEAT ? (size_t)o->n_save : 0;
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
    // what the kernels read: the number of save points per step (index 0: the u0 slot)
    std::vector<int> cnt((size_t)o->n_steps + 1, 0);
    for (int st_ : plan_step) if (st_ >= 0 && (int64_t)st_ <= o->n_steps) ++cnt[(size_t)st_];
    plan_step.swap(cnt);
  }
  auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t off_grid = 16 * kQueueSlots;
  const size_t off_save = off_grid + up16(ng * sizeof(T));
  const size_t off_pstep = off_save + up16((adaptive ? ns : 0) * sizeof(T));
  const size_t off_pb = off_pstep + up16(plan_step.size() * sizeof(int));
  const size_t bytes = off_pb + up16(plan_b.size() * sizeof(T));
  SDE_CUDA(cudaMallocFromPoolAsync((void**)&c->scratch, bytes, pool, st));
  // pageable sources are staged by the runtime before cudaMemcpyAsync returns, so the vectors may die
  if (ng) {
    SDE_CUDA(cudaMemcpyAsync(c->scratch + off_grid, tg.data(), ng * sizeof(T), cudaMemcpyHostToDevice, st));
    c->tgrid = c->scratch + off_grid;
  }
  if (adaptive && ns) {
    SDE_CUDA(cudaMemcpyAsync(c->scratch + off_save, o->saveat, ns * sizeof(T), cudaMemcpyHostToDevice, st));
    c->saveat = c->scratch + off_save;
  }
  if (!plan_step.empty()) {
    SDE_CUDA(cudaMemcpyAsync(c->scratch + off_pstep, plan_step.data(), plan_step.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    SDE_CUDA(cudaMemcpyAsync(c->scratch + off_pb, plan_b.data(), plan_b.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    c->plan_cnt = (const int*)(c->scratch + off_pstep);
    c->plan_b = c->scratch + off_pb;
  }
  return SDE_OK;
}

int upload_consts(const sde_options_t* o, cudaMemPool_t pool, cudaStream_t st, SolveConsts* c) {
  return o->dtype == SDE_F64 ? upload_consts_t<double>(o, pool, st, c) : upload_consts_t<float>(o, pool, st, c);
}

// --------------------------------------------------------------------------------------------
// launch one piece (o->n_traj trajectories) on the current device
// --------------------------------------------------------------------------------------------
template <class T>
int launch_piece_t(sde_system_s* sys, const sde_options_t* o, const void* fn, const SolveConsts& c, int qslot,
                   const void* d_u0, const void* d_p, int64_t ld_in, void* d_out_u, int64_t ld_out,
                   void* d_out_t, int32_t* d_nacc, int32_t* d_nrej, int32_t* d_ret, cudaStream_t st) {
  if (o->n_traj <= 0) return SDE_OK;
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
  // the kernels only look at kCompatRuntimeZero (bit 30), which must reach them as 0; the variants were chosen by get_kernel
  a.compat = o->compat & (SDE_COMPAT_FIX_VERN9_INTERP | SDE_COMPAT_STRICT_CONTROLLER | SDE_COMPAT_LOG2_CONTROLLER | SDE_COMPAT_FAST_RHS | SDE_COMPAT_FAST_STAGES);
  if (want_fast_stages(o)) {     // h_ij = dt * a_ij in the state's precision, like the product the reference-exact kernels form for a21
    static_assert(sde_host::kTsit5NStageCoef == (int)(sizeof(a.hcoef) / sizeof(a.hcoef[0])), "stage coefficients");
    for (int k = 0; k < sde_host::kTsit5NStageCoef; ++k) a.hcoef[k] = (T)((T)o->dt * (T)sde_host::kTsit5StageCoef[k]);
  }
  a.layout = o->layout;
  a.max_attempts = o->max_attempts;
  a.out_u = (T*)d_out_u;
  a.ld_out = ld_out;
  a.n_out = out_slots(o);
  a.out_t = adaptive ? (T*)d_out_t : nullptr;
  a.naccept = d_nacc; a.nreject = d_nrej; a.retcode = d_ret;
  a.queue = c.queue(qslot);
  a.tgrid = (const T*)c.tgrid;
  a.saveat = (const T*)c.saveat;
  a.plan_cnt = c.plan_cnt;
  a.plan_b = (const T*)c.plan_b;

  int dev = 0, sms = 0, per_sm = 0;
  SDE_CUDA(cudaGetDevice(&dev));
  unsigned grid;
  const int64_t full = (o->n_traj + kBlock - 1) / kBlock;
  size_t smem = 0;
  if (want_staged(o)) {
    smem = staged_smem_bytes(sys->n_state, sizeof(T), kBlock, !sys->builtin);
  } else if (want_row_stage(o, sys->n_state)) {
    smem = (size_t)kBlock * sde::kRowStageBytesPerThread;       // adaptive every-step rows (adaptive_body): rings + line table
  }
  if (smem > 48 * 1024)
    SDE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (adaptive) {
    SDE_CUDA(cudaMemsetAsync(a.queue, 0, 16, st));
    SDE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SDE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kBlock, smem));
    if (per_sm < 1) per_sm = 1;
    // Ensembles of about as many trajectories as there are resident lanes: with one trajectory per lane a warp issues
    // until its longest lane is done, while lanes that take a second trajectory from the queue even the warps out.
    // Aim at two trajectories per lane, but keep at least three CTAs per SM for latency hiding (B200, 2^16 / 2^17
    // trajectories: -4 ... -20 % time on sorted and shuffled Van der Pol / Lorenz sweeps, neutral from 2^18 on, where a
    // smaller grid helps shuffled sweeps and hurts sorted ones: profiles/r2_cta_cap_probe.txt).
    {
      const int64_t two_per_lane = (o->n_traj + (int64_t)sms * kBlock * 2 - 1) / ((int64_t)sms * kBlock * 2);
      per_sm = (int)std::min<int64_t>(per_sm, std::max<int64_t>(3, two_per_lane));
    }
    if (const char* e = getenv("SDE_TUNE_CTAS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e)));   // measurement only
    grid = (unsigned)std::min<int64_t>(full, (int64_t)sms * per_sm);   // persistent CTAs
  } else {
    if (full > 0x7fffffffLL) return fail(SDE_ERR_INVALID, "n_traj too large for one launch");
    grid = (unsigned)full;
    // fixed step: no step control and no failure mode -> the per-trajectory statistics are all zero
    if (d_nacc) SDE_CUDA(cudaMemsetAsync(d_nacc, 0, 4 * (size_t)o->n_traj, st));
    if (d_nrej) SDE_CUDA(cudaMemsetAsync(d_nrej, 0, 4 * (size_t)o->n_traj, st));
    if (d_ret) SDE_CUDA(cudaMemsetAsync(d_ret, 0, 4 * (size_t)o->n_traj, st));
  }
  void* params[] = {&a};
  SDE_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(kBlock), params, smem, st));
  g_launches.fetch_add(1);
  return SDE_OK;
}

int launch_piece(sde_system_s* sys, const sde_options_t* o, const void* fn, const SolveConsts& c, int qslot,
                 const void* d_u0, const void* d_p, int64_t ld_in, void* d_out_u, int64_t ld_out,
                 void* d_out_t, int32_t* d_nacc, int32_t* d_nrej, int32_t* d_ret, cudaStream_t st) {
  if (o->dtype == SDE_F64)
    return launch_piece_t<double>(sys, o, fn, c, qslot, d_u0, d_p, ld_in, d_out_u, ld_out, d_out_t, d_nacc, d_nrej, d_ret, st);
  return launch_piece_t<float>(sys, o, fn, c, qslot, d_u0, d_p, ld_in, d_out_u, ld_out, d_out_t, d_nacc, d_nrej, d_ret, st);
}

// device-resident solve: constants + one launch on the caller's stream
int launch(sde_system_s* sys, const sde_options_t* o, const void* d_u0, const void* d_p, int64_t ld_in,
           void* d_out_u, int64_t ld_out, void* d_out_t, int32_t* d_nacc, int32_t* d_nrej,
           int32_t* d_ret, cudaStream_t st) {
  const void* fn = nullptr;
  int rc = get_kernel(sys, o, true, &fn);
  if (rc != SDE_OK) return rc;
  int dev = 0;
  SDE_CUDA(cudaGetDevice(&dev));
  cudaMemPool_t pool;
  rc = device_pool(dev, &pool);
  if (rc != SDE_OK) return rc;
  SolveConsts c;
  rc = upload_consts(o, pool, st, &c);
  if (rc == SDE_OK)
    rc = launch_piece(sys, o, fn, c, 0, d_u0, d_p, ld_in, d_out_u, ld_out, d_out_t, d_nacc, d_nrej, d_ret, st);
  if (c.scratch) cudaFreeAsync(c.scratch, st);
  return rc;
}


// --------------------------------------------------------------------------------------------
// Source of contiguous trajectory ranges for one device of a host-buffer solve.
//   static : the device owns [lo, hi) = [floor(gN/G), floor((g+1)N/G)) and walks it in pieces that
//            fit its memory budget (fixed-step algorithms: every trajectory costs the same);
//   dynamic: all devices pull pieces (guided self-scheduling, >= `grain`) from one shared counter (adaptive
//            algorithms: step counts vary 10x along a parameter sweep, so equal index ranges are
//            not equal work -- the host-level analogue of the kernels' work queue).
// --------------------------------------------------------------------------------------------
struct RangeSource {
  int64_t lo = 0, hi = 0;                    // static range (cursor = lo)
  std::atomic<int64_t>* shared = nullptr;    // dynamic: next unassigned trajectory
  int64_t total = 0, grain = 0;              // dynamic: ensemble size, smallest piece handed out
  int n_dev = 1;
  // dynamic pieces follow guided self-scheduling: half of an equal share of what is left, never less
  // than `grain` -- large pieces first, small ones at the end, where the imbalance is decided
  int64_t guided_len(int64_t s) const {
    return std::max<int64_t>(grain, (((total - s) / (2 * (int64_t)n_dev)) + 31) & ~(int64_t)31);
  }
  bool next(int64_t max_len, int64_t* a, int64_t* b) {
    if (shared) {
      int64_t s = shared->load();
      int64_t len;
      do {
        if (s >= total) return false;
        len = std::min(max_len, guided_len(s));
      } while (!shared->compare_exchange_weak(s, s + len));
      *a = s; *b = std::min(total, s + len);
      return true;
    }
    if (lo >= hi) return false;
    *a = lo; *b = std::min(hi, lo + max_len);
    lo = *b;
    return true;
  }
  int64_t max_piece() const { return shared ? std::min(total, guided_len(0)) : hi - lo; }
};

// --------------------------------------------------------------------------------------------
// Small host-buffer solves (config 1 of BASELINE.json: 10 k trajectories, a 0.59 ms kernel).  The general path below
// costs ~30 runtime calls per solve -- a stream, seven stream-ordered allocations, two pitched H2D copies, up to six D2H
// copies into pageable memory (each one a synchronisation), the frees, the pool trim: 0.25 ms around that kernel
// (tools/small_solve_overhead.py).  A solve whose inputs and outputs fit kSmallSolveBytes instead borrows a cached
// context -- one stream, one device buffer, one pinned staging buffer -- packs its inputs into the staging buffer,
// and runs ONE H2D copy, the kernel and ONE D2H copy on that stream; the outputs are scattered to the caller's arrays
// by the host.  Contexts are pooled (any thread, one user at a time) and released by sde_trim().
// --------------------------------------------------------------------------------------------
constexpr size_t kSmallSolveBytes = (size_t)4 << 20;
struct SmallCtx {
  int dev = -1;
  cudaStream_t st = nullptr;
  char* d = nullptr;       // device buffer
  char* h = nullptr;       // pinned host staging, same layout
  size_t cap = 0;
};
std::mutex g_small_mu;
std::vector<SmallCtx*> g_small_free;

void small_destroy(SmallCtx* c) {
  if (c->d) cudaFree(c->d);
  if (c->h) cudaFreeHost(c->h);
  if (c->st) cudaStreamDestroy(c->st);
  delete c;
}
// a context of the current device with at least `bytes` of device and staging memory
int small_acquire(int dev, size_t bytes, SmallCtx** out) {
  SmallCtx* c = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_small_mu);
    for (size_t i = 0; i < g_small_free.size(); ++i)
      if (g_small_free[i]->dev == dev) { c = g_small_free[i]; g_small_free.erase(g_small_free.begin() + (long)i); break; }
  }
  if (!c) { c = new SmallCtx; c->dev = dev; }
  cudaError_t e = cudaSuccess;
  if (!c->st) e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking);
  if (e == cudaSuccess && c->cap < bytes) {
    if (c->d) { cudaFree(c->d); c->d = nullptr; }
    if (c->h) { cudaFreeHost(c->h); c->h = nullptr; }
    c->cap = 0;
    const size_t cap = std::max<size_t>((bytes + 0xfffff) & ~(size_t)0xfffff, (size_t)1 << 20);
    e = cudaMalloc((void**)&c->d, cap);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&c->h, cap, cudaHostAllocDefault);
    if (e == cudaSuccess) c->cap = cap;
  }
  if (e != cudaSuccess) {
    small_destroy(c);
    return fail(SDE_ERR_CUDA, "small-solve context: %s", cudaGetErrorString(e));
  }
  *out = c;
  return SDE_OK;
}
void small_release(SmallCtx* c) {
  std::lock_guard<std::mutex> lk(g_small_mu);
  g_small_free.push_back(c);
}
void small_trim() {
  std::vector<SmallCtx*> all;
  {
    std::lock_guard<std::mutex> lk(g_small_mu);
    all.swap(g_small_free);
  }
  for (SmallCtx* c : all) small_destroy(c);
}

// trajectories [c0, c1) of a host-buffer solve on the current device through a cached context (see above)
int small_solve(sde_system_s* sys, const sde_options_t* o, const void* fn, int dev, int64_t c0, int64_t c1,
                const char* u0, const char* p, char* out_u, char* out_t, int32_t* nacc, int32_t* nrej, int32_t* ret,
                bool trace) {
  const size_t es = esize(o->dtype);
  const int N = sys->n_state, NP = sys->n_param;
  const int64_t n_all = o->n_traj, n = c1 - c0, slots = out_slots(o);
  const bool series = o->save_mode != SDE_SAVE_ENDPOINT;
  const bool adaptive = is_adaptive(o->alg);
  const bool t_series = adaptive && o->save_mode == SDE_SAVE_EVERYSTEP;
  const bool tm = o->layout == SDE_LAYOUT_TRAJ_MAJOR;
  const int64_t ld = (n + 31) & ~(int64_t)31;         // device pitch of the SoA rows, in elements
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t at = off; off += up(bytes); return at; };
  const size_t o_u0 = take(es * N * ld), o_p = take(es * NP * ld);
  const size_t in_bytes = off;
  const size_t o_out = take(es * N * slots * ld);
  const size_t o_t = (adaptive && out_t) ? take(es * (t_series ? slots : 1) * ld) : 0;
  const size_t o_na = nacc ? take(4 * ld) : 0, o_nr = nrej ? take(4 * ld) : 0, o_rc = ret ? take(4 * ld) : 0;
  const size_t out_bytes = off - in_bytes;
  const size_t o_q = take(16 * kQueueSlots);          // work-queue heads (adaptive kernels)
  auto now = []() { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(now() - a).count(); };
  auto t_phase = now();
  SmallCtx* c = nullptr;
  int rc = small_acquire(dev, off, &c);
  if (rc != SDE_OK) return rc;
  cudaStream_t st = c->st;
  SolveConsts consts;
  bool consts_owned = false;
  auto finish = [&](int code) {       // nothing in flight may still use the context when it goes back to the pool
    cudaStreamSynchronize(st);
    small_release(c);
    return code;
  };
#define SDE_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
      return finish(fail(SDE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); } while (0)
  // inputs: SoA rows of the range, packed with the device pitch
  for (int r = 0; r < N; ++r) memcpy(c->h + o_u0 + es * r * ld, u0 + es * ((size_t)r * n_all + c0), es * n);
  for (int r = 0; r < NP; ++r) memcpy(c->h + o_p + es * r * ld, p + es * ((size_t)r * n_all + c0), es * n);
  SDE_TRY(cudaMemcpyAsync(c->d, c->h, in_bytes, cudaMemcpyHostToDevice, st));
  // the one D2H copy below also moves the padding between the output arrays (row pitch, 256-byte alignment), which no
  // kernel writes: defined bytes instead of whatever the buffer held (compute-sanitizer initcheck reads them as uninitialised)
  SDE_TRY(cudaMemsetAsync(c->d + in_bytes, 0, out_bytes, st));
  // constants: adaptive solves without saveat need only the queue heads, which live in the context's buffer
  if (adaptive && o->save_mode != SDE_SAVE_SAVEAT) {
    consts.scratch = c->d + o_q;
  } else {
    cudaMemPool_t pool;
    rc = device_pool(dev, &pool);
    if (rc == SDE_OK) rc = upload_consts(o, pool, st, &consts);
    if (rc != SDE_OK) return finish(rc);
    consts_owned = true;
  }
  sde_options_t oc = *o;
  oc.n_traj = n;
  rc = launch_piece(sys, &oc, fn, consts, 0, c->d + o_u0, c->d + o_p, ld, c->d + o_out, ld,
                    (adaptive && out_t) ? c->d + o_t : nullptr, nacc ? (int32_t*)(c->d + o_na) : nullptr,
                    nrej ? (int32_t*)(c->d + o_nr) : nullptr, ret ? (int32_t*)(c->d + o_rc) : nullptr, st);
  if (consts_owned && consts.scratch) cudaFreeAsync(consts.scratch, st);      // stream-ordered: after the kernel
  if (rc != SDE_OK) return finish(rc);
  SDE_TRY(cudaMemcpyAsync(c->h + in_bytes, c->d + in_bytes, out_bytes, cudaMemcpyDeviceToHost, st));
  SDE_TRY(cudaStreamSynchronize(st));
#undef SDE_TRY
  if (trace) { fprintf(stderr, "[sde trace] dev %d: [%lld,%lld) small solve, device part %.3f ms\n", dev, (long long)c0, (long long)c1, ms_since(t_phase)); t_phase = now(); }
  // outputs back to the caller's layout (the same index arithmetic as the D2H copies of the general path)
  const char* h = c->h;
  if (!series) {
    for (int r = 0; r < N; ++r) memcpy(out_u + es * ((size_t)r * n_all + c0), h + o_out + es * r * ld, es * n);
  } else if (tm) {
    memcpy(out_u + es * N * slots * c0, h + o_out, es * N * slots * n);
  } else {
    for (int64_t r = 0; r < (int64_t)N * slots; ++r) memcpy(out_u + es * ((size_t)r * n_all + c0), h + o_out + es * r * ld, es * n);
  }
  if (adaptive && out_t) {
    if (!t_series) memcpy(out_t + es * c0, h + o_t, es * n);
    else if (tm) memcpy(out_t + es * slots * c0, h + o_t, es * slots * n);
    else for (int64_t r = 0; r < slots; ++r) memcpy(out_t + es * ((size_t)r * n_all + c0), h + o_t + es * r * ld, es * n);
  }
  if (nacc) memcpy(nacc + c0, h + o_na, 4 * n);
  if (nrej) memcpy(nrej + c0, h + o_nr, 4 * n);
  if (ret) memcpy(ret + c0, h + o_rc, 4 * n);
  small_release(c);
  if (trace) fprintf(stderr, "[sde trace] dev %d: small solve, scatter + release %.3f ms\n", dev, ms_since(t_phase));
  return SDE_OK;
}

// one device of a host-buffer solve.
// The device's range is cut into pieces; piece i uses buffer set i % 2 on stream i % 2, each stream
// running H2D -> kernel -> D2H in order, so the copies of one piece overlap the kernel of its
// neighbour and the next kernel's CTAs fill the SMs the previous kernel's tail leaves idle.
int solve_shard(sde_system_s* sys, const sde_options_t* o, int device, RangeSource src,
                const char* u0, const char* p, char* out_u, char* out_t, int32_t* nacc, int32_t* nrej,
                int32_t* ret, std::string* err) {
  constexpr int kBuf = 2;
  struct Buffers {
    char *u0 = nullptr, *p = nullptr, *out = nullptr, *t = nullptr;
    int32_t *na = nullptr, *nr = nullptr, *rc = nullptr;
  };
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
    const size_t es = esize(o->dtype);
    const int N = sys->n_state, NP = sys->n_param;
    const int64_t n_all = o->n_traj, slots = out_slots(o);
    const bool series = o->save_mode != SDE_SAVE_ENDPOINT;
    const bool adaptive = is_adaptive(o->alg);
    const bool t_series = adaptive && o->save_mode == SDE_SAVE_EVERYSTEP;
    // SDE_TRACE=1: wall-clock phases of the host-buffer path on stderr (adds stream syncs)
    const bool trace = getenv("SDE_TRACE") != nullptr;
    auto now = []() { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point a) {
      return std::chrono::duration<double, std::milli>(now() - a).count();
    };
    auto t_phase = now();
    const void* fn = nullptr;
    int rc = get_kernel(sys, o, true, &fn);
    if (rc != SDE_OK) return rc;
    cudaMemPool_t pool;
    rc = device_pool(dev, &pool);
    if (rc != SDE_OK) return rc;

    // piece size: a quarter of the range (at least 2^17 trajectories, 2^21 for adaptive kernels), bounded by the memory budget
    // (60 % of what is free now plus what the pool already holds, shared by the kBuf buffer sets)
    const size_t per_traj = es * ((size_t)N + NP + (size_t)N * slots + (t_series ? slots : 1)) + 12;
    const int64_t range = src.max_piece();
    unsigned long long pooled = 0;
    int64_t piece = range;
    if ((double)per_traj * (double)range * kBuf > 512.0 * 1048576.0) {
      // large solves: bound the pieces by what is free now plus the idle bytes the pool already holds
      // (cudaMemGetInfo costs ~0.3 ms, so small solves skip it: any B200 has 512 MB to spare)
      size_t free_b = 0, total_b = 0;
      SDE_CUDA(cudaMemGetInfo(&free_b, &total_b));
      unsigned long long pool_used = 0;
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &pooled);
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &pool_used);
      pooled = pooled > pool_used ? pooled - pool_used : 0;
      piece = (int64_t)std::max<size_t>(1, (size_t)(0.6 * (double)(free_b + pooled)) / per_traj / kBuf);
    }
    // adaptive kernels are persistent (~1.5e5 lanes in flight) and end with a latency-bound tail of the
    // longest trajectories, so their pieces must hold many trajectories per lane to stay efficient
    int64_t want = std::max<int64_t>((range + 3) / 4, (int64_t)1 << (adaptive ? 21 : 17));
    if (const char* e = getenv("SDE_TUNE_PIECE")) want = std::max<int64_t>(32, atoll(e));   // measurement only
    piece = std::min<int64_t>(piece, std::min<int64_t>(want, range));
    if (piece > 32) piece = (piece + 31) & ~(int64_t)31;
    int n_buf = piece >= range ? 1 : kBuf;
    if (const char* e = getenv("SDE_TUNE_NBUF")) n_buf = std::max(1, std::min(kBuf, atoi(e)));   // measurement only

    // small solves: one cached context, one copy each way (small_solve above); SDE_TUNE_NO_SMALL = measurement only
    if (!src.shared && n_buf == 1 && piece >= range && range > 0 && !getenv("SDE_TUNE_NO_SMALL") && !getenv("SDE_TUNE_PIECE") &&
        per_traj * (size_t)range <= kSmallSolveBytes) {
      int64_t c0 = 0, c1 = 0;
      src.next(range, &c0, &c1);
      return small_solve(sys, o, fn, dev, c0, c1, u0, p, out_u, out_t, nacc, nrej, ret, trace);
    }

    cudaStream_t st[kBuf] = {nullptr, nullptr};
    Buffers buf[kBuf];
    SolveConsts consts;
    auto cleanup = [&]() {
      for (int b = 0; b < kBuf; ++b) if (st[b]) cudaStreamSynchronize(st[b]);   // nothing in flight uses the buffers
      for (int b = 0; b < kBuf; ++b) {
        if (!st[b]) continue;
        Buffers& B = buf[b];
        void* ptrs[] = {B.u0, B.p, B.out, B.t, B.na, B.nr, B.rc};
        for (void* q : ptrs) if (q) cudaFreeAsync(q, st[b]);
      }
      if (consts.scratch && st[0]) cudaFreeAsync(consts.scratch, st[0]);
      for (int b = 0; b < kBuf; ++b) if (st[b]) { cudaStreamSynchronize(st[b]); cudaStreamDestroy(st[b]); }
      cudaMemPoolTrimTo(pool, pool_keep_bytes());
    };
#define SDE_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); \
      return fail(SDE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
    if (piece <= 0) return SDE_OK;
    for (int b = 0; b < n_buf; ++b) {
      SDE_TRY(cudaStreamCreateWithFlags(&st[b], cudaStreamNonBlocking));
      Buffers& B = buf[b];
      SDE_TRY(cudaMallocFromPoolAsync((void**)&B.u0, es * N * piece, pool, st[b]));
      if (NP) SDE_TRY(cudaMallocFromPoolAsync((void**)&B.p, es * NP * piece, pool, st[b]));
      SDE_TRY(cudaMallocFromPoolAsync((void**)&B.out, es * N * slots * piece, pool, st[b]));
      if (adaptive && out_t) SDE_TRY(cudaMallocFromPoolAsync((void**)&B.t, es * piece * (t_series ? slots : 1), pool, st[b]));
      if (nacc) SDE_TRY(cudaMallocFromPoolAsync((void**)&B.na, 4 * piece, pool, st[b]));
      if (nrej) SDE_TRY(cudaMallocFromPoolAsync((void**)&B.nr, 4 * piece, pool, st[b]));
      if (ret) SDE_TRY(cudaMallocFromPoolAsync((void**)&B.rc, 4 * piece, pool, st[b]));
    }
    // constants once, on stream 0; the other stream waits for them
    rc = upload_consts(o, pool, st[0], &consts);
    if (rc != SDE_OK) { cleanup(); return rc; }
    if (n_buf > 1) {
      cudaEvent_t ready;
      SDE_TRY(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
      cudaEventRecord(ready, st[0]);
      cudaStreamWaitEvent(st[1], ready, 0);
      cudaEventDestroy(ready);
    }
    if (trace) {
      for (int b = 0; b < n_buf; ++b) cudaStreamSynchronize(st[b]);
      fprintf(stderr, "[sde trace] dev %d: setup (piece %lld x %d buffers, pool held %.1f MB) %.3f ms\n", dev,
              (long long)piece, n_buf, (double)pooled / 1048576.0, ms_since(t_phase));
    }
    int64_t c0 = 0, c1 = 0;
    for (int i = 0;; ++i) {
      const int b = i % n_buf;
      // dynamic ranges: a device asks for more work only when one of its buffer sets is free again --
      // launches are asynchronous, so without this the first host thread to finish its setup would
      // drain the shared cursor before the other devices have created their streams
      if (src.shared) SDE_TRY(cudaStreamSynchronize(st[b]));
      if (!src.next(piece, &c0, &c1)) break;
      const Buffers& B = buf[b];
      cudaStream_t s_ = st[b];
      const int64_t n = c1 - c0;
      t_phase = now();
      // H2D: SoA rows of the piece (host pitch = n_all elements, device pitch = piece elements)
      SDE_TRY(cudaMemcpy2DAsync(B.u0, es * piece, u0 + es * c0, es * n_all, es * n, N, cudaMemcpyHostToDevice, s_));
      if (NP) SDE_TRY(cudaMemcpy2DAsync(B.p, es * piece, p + es * c0, es * n_all, es * n, NP, cudaMemcpyHostToDevice, s_));
      if (trace) { cudaStreamSynchronize(s_); fprintf(stderr, "[sde trace] dev %d: [%lld,%lld) H2D %.3f ms\n", dev, (long long)c0, (long long)c1, ms_since(t_phase)); t_phase = now(); }
      sde_options_t oc = *o;
      oc.n_traj = n;
      rc = launch_piece(sys, &oc, fn, consts, b, B.u0, B.p, piece, B.out, piece, B.t, B.na, B.nr, B.rc, s_);
      if (rc != SDE_OK) { cleanup(); return rc; }
      if (trace) { cudaStreamSynchronize(s_); fprintf(stderr, "[sde trace] dev %d: [%lld,%lld) kernel %.3f ms\n", dev, (long long)c0, (long long)c1, ms_since(t_phase)); t_phase = now(); }
      // D2H
      if (!series) {
        SDE_TRY(cudaMemcpy2DAsync(out_u + es * c0, es * n_all, B.out, es * piece, es * n, N, cudaMemcpyDeviceToHost, s_));
      } else if (o->layout == SDE_LAYOUT_TRAJ_MAJOR) {
        SDE_TRY(cudaMemcpyAsync(out_u + es * N * slots * c0, B.out, es * N * slots * n, cudaMemcpyDeviceToHost, s_));
      } else {
        SDE_TRY(cudaMemcpy2DAsync(out_u + es * c0, es * n_all, B.out, es * piece, es * n, (size_t)N * slots, cudaMemcpyDeviceToHost, s_));
      }
      if (B.t) {
        if (!t_series) SDE_TRY(cudaMemcpyAsync(out_t + es * c0, B.t, es * n, cudaMemcpyDeviceToHost, s_));
        else if (o->layout == SDE_LAYOUT_TRAJ_MAJOR)
          SDE_TRY(cudaMemcpyAsync(out_t + es * slots * c0, B.t, es * slots * n, cudaMemcpyDeviceToHost, s_));
        else
          SDE_TRY(cudaMemcpy2DAsync(out_t + es * c0, es * n_all, B.t, es * piece, es * n, (size_t)slots, cudaMemcpyDeviceToHost, s_));
      }
      if (B.na) SDE_TRY(cudaMemcpyAsync(nacc + c0, B.na, 4 * n, cudaMemcpyDeviceToHost, s_));
      if (B.nr) SDE_TRY(cudaMemcpyAsync(nrej + c0, B.nr, 4 * n, cudaMemcpyDeviceToHost, s_));
      if (B.rc) SDE_TRY(cudaMemcpyAsync(ret + c0, B.rc, 4 * n, cudaMemcpyDeviceToHost, s_));
      if (trace) { cudaStreamSynchronize(s_); fprintf(stderr, "[sde trace] dev %d: [%lld,%lld) D2H %.3f ms\n", dev, (long long)c0, (long long)c1, ms_since(t_phase)); }
    }
    for (int b = 0; b < n_buf; ++b) SDE_TRY(cudaStreamSynchronize(st[b]));
#undef SDE_TRY
    t_phase = now();
    cleanup();
    if (trace) fprintf(stderr, "[sde trace] dev %d: release + trim %.3f ms\n", dev, ms_since(t_phase));
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
    int rc = nvrtc_compile(user_program(s, 0, dtype, 0, false, false, false, true), nullptr, &lg, false);
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
  if (opt->n_traj == 0) return SDE_OK;
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
  if (n_dev < 0 || (n_dev > 0 && !devices)) return fail(SDE_ERR_INVALID, "bad device list");
  if (opt->n_traj == 0) return SDE_OK;      // an empty ensemble: nothing to read or write
  if (!u0 || !out_u || (sys->n_param > 0 && !p)) return fail(SDE_ERR_INVALID, "null host buffer");
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
      src.n_dev = n_dev;
      src.grain = 1 << 19;                     // never launches too small to fill a GPU (~1.5e5 lanes in flight)
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

int sde_trim(void) {
  small_trim();
  std::lock_guard<std::mutex> lk(g_pool_mu);
  for (auto& kv : g_pools) {
    cudaError_t e = cudaMemPoolTrimTo(kv.second, 0);
    if (e != cudaSuccess) return fail(SDE_ERR_CUDA, "cudaMemPoolTrimTo: %s", cudaGetErrorString(e));
  }
  return SDE_OK;
}

}  // extern "C"
