// Run-time compilation: NVRTC (dlopen) -> cubin -> driver-API module per device. See ebn_jit.h.
#include "ebn_jit.h"

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <sys/stat.h>
#include <unistd.h>

#include <cuda.h>
#include <dlfcn.h>

#include "../../include/ebncuda.h"
#include "ebn_device.cuh"
#include "ebn_expr.h"
#include "ebn_jit_sources.inc"

namespace ebn {

struct JitKernel {
  std::string cubin;
  std::string entry;                 // mangled kernel name
  std::map<int, CUmodule> modules;   // per device ordinal
  std::map<int, CUfunction> funcs;
};

namespace {

std::string fmt(const char* f, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof buf, f, ap);
  va_end(ap);
  return buf;
}

// ---- NVRTC through dlopen ------------------------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct NvrtcApi {
  void* handle = nullptr;
  std::string why;  // why loading failed
  int (*Version)(int*, int*) = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
  int (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int major = 0, minor = 0;
} g_rtc;
std::mutex g_mu;

bool nvrtc_load() {
  if (g_rtc.handle) return true;
  if (!g_rtc.why.empty()) return false;
  std::vector<std::string> names;
  if (const char* e = getenv("EBN_NVRTC_LIB")) names.push_back(e);
  names.insert(names.end(), {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                             "/usr/local/cuda/lib64/libnvrtc.so"});
  void* h = nullptr;
  std::string errs;
  for (const auto& n : names) {
    h = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (h) break;
    errs += std::string(errs.empty() ? "" : "; ") + dlerror();
  }
  if (!h) {
    g_rtc.why = "cannot dlopen libnvrtc (set EBN_NVRTC_LIB): " + errs;
    return false;
  }
#define SYM(field, name)                                     \
  *(void**)(&g_rtc.field) = dlsym(h, name);                  \
  if (!g_rtc.field) {                                        \
    g_rtc.why = std::string("NVRTC symbol not found: ") + name; \
    return false;                                            \
  }
  SYM(Version, "nvrtcVersion")
  SYM(CreateProgram, "nvrtcCreateProgram")
  SYM(DestroyProgram, "nvrtcDestroyProgram")
  SYM(CompileProgram, "nvrtcCompileProgram")
  SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
  SYM(GetProgramLog, "nvrtcGetProgramLog")
  SYM(GetCUBINSize, "nvrtcGetCUBINSize")
  SYM(GetCUBIN, "nvrtcGetCUBIN")
  SYM(AddNameExpression, "nvrtcAddNameExpression")
  SYM(GetLoweredName, "nvrtcGetLoweredName")
  SYM(GetErrorString, "nvrtcGetErrorString")
#undef SYM
  g_rtc.Version(&g_rtc.major, &g_rtc.minor);
  g_rtc.handle = h;
  return true;
}

// ---- driver API through the runtime (no libcuda link dependency) ----------------------------
struct DriverApi {
  bool ready = false;
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                           CUstream, void**, void**) = nullptr;
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
} g_drv;

bool driver_load(std::string* err) {
  if (g_drv.ready) return true;
  struct { void** p; const char* name; } syms[] = {
      {(void**)&g_drv.ModuleLoadData, "cuModuleLoadData"},
      {(void**)&g_drv.ModuleUnload, "cuModuleUnload"},
      {(void**)&g_drv.ModuleGetFunction, "cuModuleGetFunction"},
      {(void**)&g_drv.LaunchKernel, "cuLaunchKernel"},
      {(void**)&g_drv.OccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor"},
      {(void**)&g_drv.GetErrorString, "cuGetErrorString"},
      {(void**)&g_drv.FuncSetAttribute, "cuFuncSetAttribute"},
  };
  for (auto& s : syms) {
    cudaDriverEntryPointQueryResult q;
    const cudaError_t e = cudaGetDriverEntryPointByVersion(s.name, s.p, 12000, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*s.p) {
      cudaGetLastError();
      *err = fmt("driver entry point %s unavailable (%s)", s.name, cudaGetErrorString(e));
      return false;
    }
  }
  g_drv.ready = true;
  return true;
}

std::string cu_err(CUresult r) {
  const char* s = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
  return s ? s : fmt("CUresult %d", (int)r);
}

// ---- source generation ----------------------------------------------------------------------
const char* ls_type(int id) {
  switch (id) {
    case EBN_LS_VEHICLE_SUSPENSION: return "VehicleSuspension";
    case EBN_LS_STRAUB_FRAME: return "StraubFrame";
    case EBN_LS_HYDROGEN_RADIUS: return "HydrogenRadius";
    case EBN_LS_POLYNOMIAL: return "PolynomialProgram";
    case EBN_LS_MIN_AFFINE: return "MinAffine";
    case EBN_LS_STRAUB_FRAME_R45: return "StraubFrameR45";
    case EBN_LS_STRAUB_CAPACITY: return "StraubCapacity";
    case EBN_LS_EXPRESSION: return "JitExpr";
  }
  return nullptr;
}

struct Unit {
  std::string source, name_expr;
};

bool make_unit(const JitSpec& spec, JitRole role, Unit* u, std::string* err) {
  const char* ls = ls_type(spec.limit_state);
  if (!ls) {
    *err = fmt("limit state %d has no device functor", spec.limit_state);
    return false;
  }
  u->source = "#include \"ebn_mc_impl.cuh\"\n";
  if (spec.limit_state == EBN_LS_EXPRESSION) {
    const std::string bad = expr_check(spec.consts, spec.n_consts, spec.d, nullptr);
    if (!bad.empty()) {
      *err = bad;
      return false;
    }
    u->source += expr_codegen(spec.consts, spec.n_consts, spec.d);
  }
  const std::string functor = fmt("ebn::%s<ebn::%s>", ls, spec.strict ? "StrictOps" : "FastOps");
  std::string plan;
  if (spec.qmc) plan = "ebn::QmcPlan";
  else if (spec.kinds.empty()) plan = spec.copula ? "ebn::CopulaPlan" : "ebn::DynamicPlan";
  else {
    plan = spec.copula ? "ebn::StaticCopulaPlan<" : "ebn::StaticPlan<";
    for (size_t j = 0; j < spec.kinds.size(); ++j) plan += fmt(j ? ",%d" : "%d", spec.kinds[j]);
    plan += ">";
  }
  switch (role) {
    case JIT_COUNT: u->name_expr = "ebn::mc_kernel<" + functor + ", " + plan + ", 0>"; break;
    case JIT_HIST: u->name_expr = "ebn::mc_kernel<" + functor + ", " + plan + ", 1>"; break;
    case JIT_REPLAY: u->name_expr = "ebn::replay_kernel<" + functor + ", " + plan + ">"; break;
    case JIT_MATRIX: u->name_expr = "ebn::matrix_kernel<" + functor + ">"; break;
  }
  return true;
}

uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull) {
  for (unsigned char c : s) {
    h ^= c;
    h *= 1099511628211ull;
  }
  return h;
}

// ---- caches ------------------------------------------------------------------------------------
std::map<std::string, std::unique_ptr<JitKernel>> g_cache;

// Compiled kernels persist across processes only when the caller names a directory (EBN_JIT_CACHE):
// the library writes nothing to disk on its own.
std::string disk_dir() {
  const char* e = getenv("EBN_JIT_CACHE");
  return e ? e : "";
}

bool disk_read(const std::string& path, JitKernel* k) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  uint64_t hdr[3];  // magic, entry length, cubin length
  bool ok = fread(hdr, 8, 3, f) == 3 && hdr[0] == 0x3142554345424eull && hdr[1] < 4096 && hdr[2] < (1u << 28);
  if (ok) {
    k->entry.resize(hdr[1]);
    k->cubin.resize(hdr[2]);
    ok = fread(&k->entry[0], 1, hdr[1], f) == hdr[1] && fread(&k->cubin[0], 1, hdr[2], f) == hdr[2];
  }
  fclose(f);
  return ok;
}

void disk_write(const std::string& dir, const std::string& path, const JitKernel& k) {
  mkdir(dir.c_str(), 0755);  // best effort; a failing cache is not an error
  const std::string tmp = path + fmt(".%d.tmp", (int)getpid());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  const uint64_t hdr[3] = {0x3142554345424eull, k.entry.size(), k.cubin.size()};
  const bool ok = fwrite(hdr, 8, 3, f) == 3 && fwrite(k.entry.data(), 1, k.entry.size(), f) == k.entry.size() &&
                  fwrite(k.cubin.data(), 1, k.cubin.size(), f) == k.cubin.size();
  fclose(f);
  if (ok) rename(tmp.c_str(), path.c_str());
  else unlink(tmp.c_str());
}

// The tuning macros this library was built with (Makefile EXTRA=...): a kernel compiled at run time must
// see the same values, or it would be launched with a block shape it was not compiled for. Macros that
// were not given on the command line keep the defaults of the embedded headers, which are this build's.
#define EBN_STR2(x) #x
#define EBN_STR(x) EBN_STR2(x)
const char* const kBuildDefines[] = {
    "-DEBN_THREADS=" EBN_STR(EBN_THREADS),
#ifdef EBN_UNROLL
    "-DEBN_UNROLL=" EBN_STR(EBN_UNROLL),
#endif
#ifdef EBN_PINGPONG
    "-DEBN_PINGPONG=" EBN_STR(EBN_PINGPONG),
#endif
#ifdef EBN_VEHICLE_BLOCKS
    "-DEBN_VEHICLE_BLOCKS=" EBN_STR(EBN_VEHICLE_BLOCKS),
#endif
#ifdef EBN_STRAUB_BLOCKS
    "-DEBN_STRAUB_BLOCKS=" EBN_STR(EBN_STRAUB_BLOCKS),
#endif
#ifdef EBN_BM_OPT
    "-DEBN_BM_OPT=" EBN_STR(EBN_BM_OPT),
#endif
#ifdef EBN_GENERIC_WARPS
    "-DEBN_GENERIC_WARPS=" EBN_STR(EBN_GENERIC_WARPS),
#endif
#ifdef EBN_PIPE_MAX_INPUTS
    "-DEBN_PIPE_MAX_INPUTS=" EBN_STR(EBN_PIPE_MAX_INPUTS),
#endif
#ifdef EBN_PHILOX_ROUNDS
    "-DEBN_PHILOX_ROUNDS=" EBN_STR(EBN_PHILOX_ROUNDS),
#endif
};
constexpr int kNumBuildDefines = sizeof(kBuildDefines) / sizeof(kBuildDefines[0]);

std::string build_defines_key() {
  std::string s;
  for (const char* d : kBuildDefines) s += std::string(d) + " ";
  return s;
}

bool compile(const Unit& u, JitKernel* k, std::string* err) {
  constexpr int NH = sizeof(kJitSources) / sizeof(kJitSources[0]);
  const char* names[NH];
  const char* texts[NH];
  for (int i = 0; i < NH; ++i) {
    names[i] = kJitSources[i].name;
    texts[i] = kJitSources[i].text;
  }
  nvrtcProgram prog = nullptr;
  int r = g_rtc.CreateProgram(&prog, u.source.c_str(), "ebn_jit_unit.cu", NH, texts, names);
  if (r) {
    *err = std::string("nvrtcCreateProgram: ") + g_rtc.GetErrorString(r);
    return false;
  }
  bool ok = false;
  do {
    if ((r = g_rtc.AddNameExpression(prog, u.name_expr.c_str()))) break;
    const char* opts[3 + kNumBuildDefines] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
    for (int i = 0; i < kNumBuildDefines; ++i) opts[3 + i] = kBuildDefines[i];
    r = g_rtc.CompileProgram(prog, 3 + kNumBuildDefines, opts);
    if (r) {
      size_t n = 0;
      g_rtc.GetProgramLogSize(prog, &n);
      std::string log(n, '\0');
      if (n) g_rtc.GetProgramLog(prog, &log[0]);
      if (log.size() > 3000) log.resize(3000);
      *err = std::string("run-time compilation failed: ") + g_rtc.GetErrorString(r) + "\n" + log;
      r = 0;
      break;
    }
    const char* lowered = nullptr;
    if ((r = g_rtc.GetLoweredName(prog, u.name_expr.c_str(), &lowered))) break;
    k->entry = lowered;
    size_t n = 0;
    if ((r = g_rtc.GetCUBINSize(prog, &n))) break;
    k->cubin.resize(n);
    if ((r = g_rtc.GetCUBIN(prog, &k->cubin[0]))) break;
    ok = true;
  } while (0);
  if (r) *err = std::string("NVRTC: ") + g_rtc.GetErrorString(r);
  g_rtc.DestroyProgram(&prog);
  return ok;
}

}  // namespace

// One-off probe: can this NVRTC generate sm_100a code at all? (A libnvrtc older than 12.8 loads fine and
// then rejects the architecture; torch bundles its own copy, so the first libnvrtc.so.12 on the search
// path is not necessarily the toolkit's.) Without the probe the OPTIONAL uses of run-time compilation
// (sampling plans of large jobs, rewritten coefficient-table functors) would fail where the compiled-in
// kernels could have run the job.
bool nvrtc_probe() {
  static int state = 0;  // 0 unknown, 1 ok, -1 failed
  if (state) return state > 0;
  nvrtcProgram prog = nullptr;
  const char* src = "extern \"C\" __global__ void ebn_probe(double* p) { p[0] = 1.0; }\n";
  int r = g_rtc.CreateProgram(&prog, src, "ebn_probe.cu", 0, nullptr, nullptr);
  if (!r) {
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17"};
    r = g_rtc.CompileProgram(prog, 2, opts);
    size_t n = 0;
    if (!r) r = g_rtc.GetCUBINSize(prog, &n);
    if (!r && n == 0) r = -1;
    g_rtc.DestroyProgram(&prog);
  }
  if (r) {
    g_rtc.why = fmt("libnvrtc %d.%d cannot compile for sm_100a (%s); point EBN_NVRTC_LIB at a CUDA >= 12.8 copy",
                    g_rtc.major, g_rtc.minor, r > 0 ? g_rtc.GetErrorString(r) : "empty cubin");
    state = -1;
    return false;
  }
  state = 1;
  return true;
}

bool jit_available(std::string* why) {
  std::lock_guard<std::mutex> lk(g_mu);
  const bool ok = nvrtc_load() && nvrtc_probe();
  if (!ok && why) *why = g_rtc.why;
  return ok;
}

bool jit_get(const JitSpec& spec, JitRole role, JitKernel** out, bool* compiled, std::string* err) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (compiled) *compiled = false;
  Unit u;
  if (!make_unit(spec, role, &u, err)) return false;
  // everything that determines the machine code: the embedded headers (a cubin cached on disk by an older
  // library must not be picked up by a newer one — McParams and the stream layout live there), the build's
  // -D options, the kernel instantiation and the generated functor
  static const std::string headers_key = [] {
    uint64_t h = 1469598103934665603ull;
    for (const auto& src : kJitSources) h = fnv1a(src.text, fnv1a(src.name, h));
    return fmt("headers:%016llx", (unsigned long long)h);
  }();
  const std::string key = headers_key + "\n" + build_defines_key() + "\n" + u.name_expr + "\n" + u.source;
  auto it = g_cache.find(key);
  if (it != g_cache.end()) {
    *out = it->second.get();
    return true;
  }
  if (!nvrtc_load() || !nvrtc_probe()) {
    *err = g_rtc.why + " — expression limit states and run-time sampling plans need it (there is no CPU fallback)";
    return false;
  }
  std::unique_ptr<JitKernel> k(new JitKernel);
  // on-disk cache: keyed by everything that determines the code
  uint64_t h = fnv1a(key);
  for (const auto& s : kJitSources) h = fnv1a(s.text, h);
  h = fnv1a(fmt("nvrtc %d.%d sm_100a", g_rtc.major, g_rtc.minor), h);
  const std::string dir = disk_dir();
  const std::string path = dir.empty() ? "" : dir + fmt("/%016llx.ebnk", (unsigned long long)h);
  if (path.empty() || !disk_read(path, k.get())) {
    if (!compile(u, k.get(), err)) return false;
    if (compiled) *compiled = true;
    if (!path.empty()) disk_write(dir, path, *k);
  }
  *out = k.get();
  g_cache[key] = std::move(k);
  return true;
}

size_t jit_cubin_bytes(const JitKernel* k) { return k ? k->cubin.size() : 0; }

namespace {
bool function_on(JitKernel* k, int device, CUfunction* fn, std::string* err) {
  auto it = k->funcs.find(device);
  if (it != k->funcs.end()) {
    *fn = it->second;
    return true;
  }
  if (!driver_load(err)) return false;
  CUmodule mod = nullptr;
  CUresult r = g_drv.ModuleLoadData(&mod, k->cubin.data());
  if (r != CUDA_SUCCESS) {
    *err = "cuModuleLoadData: " + cu_err(r);
    return false;
  }
  r = g_drv.ModuleGetFunction(fn, mod, k->entry.c_str());
  if (r != CUDA_SUCCESS) {
    g_drv.ModuleUnload(mod);
    *err = "cuModuleGetFunction(" + k->entry + "): " + cu_err(r);
    return false;
  }
  k->modules[device] = mod;
  k->funcs[device] = *fn;
  return true;
}
}  // namespace

bool jit_occupancy(JitKernel* k, int device, int threads, size_t dyn_smem, int* blocks_per_sm, std::string* err) {
  std::lock_guard<std::mutex> lk(g_mu);
  CUfunction fn;
  if (!function_on(k, device, &fn, err)) return false;
  // static + dynamic shared memory may pass 48 KB (histogram edges and bins next to the tables): opt in
  if (dyn_smem) g_drv.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dyn_smem);
  const CUresult r = g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fn, threads, dyn_smem);
  if (r != CUDA_SUCCESS) {
    *err = "cuOccupancyMaxActiveBlocksPerMultiprocessor: " + cu_err(r);
    return false;
  }
  return true;
}

bool jit_launch(JitKernel* k, int device, int blocks, int threads, size_t dyn_smem, cudaStream_t st, void** args,
                std::string* err) {
  std::lock_guard<std::mutex> lk(g_mu);
  CUfunction fn;
  if (!function_on(k, device, &fn, err)) return false;
  if (dyn_smem) g_drv.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dyn_smem);
  const CUresult r = g_drv.LaunchKernel(fn, (unsigned)blocks, 1, 1, (unsigned)threads, 1, 1, (unsigned)dyn_smem,
                                        (CUstream)st, args, nullptr);
  if (r != CUDA_SUCCESS) {
    *err = "cuLaunchKernel(" + k->entry + "): " + cu_err(r);
    return false;
  }
  return true;
}

void jit_unload_all() {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& kv : g_cache) {
    for (auto& m : kv.second->modules)
      if (g_drv.ModuleUnload && cudaSetDevice(m.first) == cudaSuccess) g_drv.ModuleUnload(m.second);
    kv.second->modules.clear();
    kv.second->funcs.clear();
  }
}

}  // namespace ebn
