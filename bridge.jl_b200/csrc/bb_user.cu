/*
 * bb_user.cu -- user-defined target processes: the reference's extension point (Bridge.b / Bridge.σ methods for an own
 * struct, src/types.jl:23,32-33, src/Bridge.jl:105-106) as CUDA C source compiled at run time.
 *
 * bb_user_model_create stores the user's drift statements and σ entries; the first launch that needs a kernel for a
 * (kind, guide, auxiliary mode, RNG mode) combination builds
 *     struct MUser { D, DP, b(), col(), sig() };   extern "C" __global__ bb_user_kernel(args) { bb_chain<MUser,...>::run(args); }
 * on top of the library's own kernel headers (embedded at build time: bb_embedded.inc) and compiles it with NVRTC for
 * sm_100a with -fmad=false -- the flags of the static build, so a user model that restates a registry model gives the
 * same bits (tests/test_gpu_user.py).  NVRTC is bound with dlopen; the cubin is loaded through the runtime's library
 * API (cudaLibraryLoadData / cudaLibraryGetKernel) and launched with cudaLaunchKernel.
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "bb_host.h"
#include "bb_embedded.inc"

namespace {
typedef struct _nvrtcProgram* nvrtc_program;
struct nvrtc_api {
  void* handle = nullptr;
  int (*CreateProgram)(nvrtc_program*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtc_program, int, const char* const*) = nullptr;
  int (*GetProgramLogSize)(nvrtc_program, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtc_program, char*) = nullptr;
  int (*GetCUBINSize)(nvrtc_program, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtc_program, char*) = nullptr;
  int (*DestroyProgram)(nvrtc_program*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
nvrtc_api g_rtc;
std::mutex g_mu;

int rtc_load(std::string& err) {
  if (g_rtc.handle) return BB_OK;
  /* the toolkit's own NVRTC first (by path): a process that has imported torch already holds torch's bundled
   * libnvrtc.so.12 (CUDA 12.8), which a dlopen by soname would return -- and its PTX back end rejects the 256-bit
   * vector loads / stores the sm_100a kernels use.  Candidates older than 12.9 are skipped. */
  const char* names[] = {getenv("BB_NVRTC_LIB"), "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so",
                         "libnvrtc.so.12", "libnvrtc.so"};
  void* h = nullptr;
  std::string tried;
  for (const char* n : names) {
    if (!n || !*n) continue;
    void* cand = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (!cand) { tried += std::string(n) + ": " + dlerror() + "; "; continue; }
    int (*ver)(int*, int*) = (int (*)(int*, int*))dlsym(cand, "nvrtcVersion");
    int major = 0, minor = 0;
    if (ver && ver(&major, &minor) == 0 && (major > 12 || (major == 12 && minor >= 9))) { h = cand; break; }
    tried += std::string(n) + ": NVRTC " + std::to_string(major) + "." + std::to_string(minor) + " < 12.9; ";
    dlclose(cand);
  }
  if (!h) {
    err = std::string("no usable NVRTC (>= 12.9): ") + tried;
    return BB_ERR_UNSUPPORTED;
  }
  nvrtc_api a;
  a.handle = h;
#define BB_SYM(field, name) a.field = (decltype(a.field))dlsym(h, name)
  BB_SYM(CreateProgram, "nvrtcCreateProgram");
  BB_SYM(CompileProgram, "nvrtcCompileProgram");
  BB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
  BB_SYM(GetProgramLog, "nvrtcGetProgramLog");
  BB_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
  BB_SYM(GetCUBIN, "nvrtcGetCUBIN");
  BB_SYM(DestroyProgram, "nvrtcDestroyProgram");
  BB_SYM(GetErrorString, "nvrtcGetErrorString");
#undef BB_SYM
  if (!a.CreateProgram || !a.CompileProgram || !a.GetCUBINSize || !a.GetCUBIN || !a.DestroyProgram) {
    err = "NVRTC library lacks a required symbol";
    return BB_ERR_UNSUPPORTED;
  }
  g_rtc = a;
  return BB_OK;
}

/* stand-ins for the system headers the kernel headers name (NVRTC has no host headers) */
const char* const k_stdint =
    "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
    "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;\n";

struct kernel_key {
  int kind, gk, gm, auxm, rng; /* kind 0: path kernel, 1: second pass (rng = mode 0/1), 2: StochasticHeun */
  bool operator<(const kernel_key& o) const {
    return memcmp(this, &o, sizeof(kernel_key)) < 0;
  }
};
struct jit_kernel {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kern = nullptr;
  size_t smem = 0;
  bool attr_set = false;
};
}  // namespace

struct bb_user_model {
  bb_ctx* ctx = nullptr;
  int d = 0, dp = 0, handle = 0;
  std::string drift, log;
  std::vector<int> col;
  std::vector<std::string> sig;
  std::map<kernel_key, jit_kernel> kernels;
};

namespace {
std::vector<bb_user_model*> g_models(1, nullptr); /* handle -> model; handle 0 is "none" */

std::string user_source(const bb_user_model* um, const kernel_key& k) {
  char buf[512];
  std::string s = "#include \"bb_second.cuh\"\n";
  s += "__device__ __forceinline__ void bb_user_b(const double* __restrict__ par, const double* x, double* o) {\n";
  s += um->drift;
  s += "\n}\nstruct MUser {\n";
  snprintf(buf, sizeof(buf), "  static constexpr int D = %d, DP = %d, ID = BB_MODEL_USER;\n  static constexpr bool SPARSE = true;\n",
           um->d, um->dp);
  s += buf;
  s += "  __device__ static __forceinline__ void b(const bb_model_dev& m, const double* x, double* o) { bb_user_b(m.par, x, o); }\n";
  s += "  __device__ static __forceinline__ constexpr int col(int i) { return ";
  for (int i = 0; i < um->d; i++) {
    snprintf(buf, sizeof(buf), "i == %d ? %d : ", i, um->col[i]);
    s += buf;
  }
  s += "-1; }\n";
  s += "  __device__ static __forceinline__ double sig(const bb_model_dev& m, int i) {\n    const double* par = m.par; (void)par;\n";
  for (int i = 0; i < um->d; i++)
    if (um->col[i] >= 0) {
      snprintf(buf, sizeof(buf), "    if (i == %d) return (double)(", i);
      s += buf;
      s += um->sig[i];
      s += ");\n";
    }
  s += "    return 0.0;\n  }\n};\n";
  const int minb = um->dp >= 2 ? 1 : BB_MINB;
  if (k.kind == 0) {
    snprintf(buf, sizeof(buf),
             "extern \"C\" __global__ void __launch_bounds__(BB_THREADS, %d) bb_user_kernel(const __grid_constant__ bb_chain_args a) {\n"
             "  bb_chain<MUser, %d, %d, %d, %d>::run(a);\n}\n",
             minb, k.gk, k.gm, k.auxm, k.rng);
    s += buf;
  } else {
    /* the second-pass kernels are __global__ templates: a wrapper cannot call them, so their bodies are instantiated
     * through an explicit instantiation and found by their lowered name -- simpler: compile a thin __global__ that
     * forwards to the same device code via the bb_second_body / bb_heun_body functions */
    if (k.kind == 1)
      snprintf(buf, sizeof(buf),
               "extern \"C\" __global__ void __launch_bounds__(BB_THREADS) bb_user_kernel(const __grid_constant__ bb_chain_args a) {\n"
               "  bb_second_body<MUser, %d, %d, %d, %d>(a);\n}\n",
               k.gk, k.gm, k.auxm, k.rng);
    else
      snprintf(buf, sizeof(buf),
               "extern \"C\" __global__ void __launch_bounds__(BB_THREADS) bb_user_kernel(const __grid_constant__ bb_chain_args a) {\n"
               "  bb_heun_body<MUser>(a);\n}\n");
    s += buf;
  }
  return s;
}

size_t path_smem(const bb_user_model* um, const kernel_key& k) {
  /* bb_chain_smem, for a run-time model */
  const int rec = bb_rec_len(k.gk, um->d, k.gm, k.auxm);
  const bool xbuf = BB_XFLUSH && um->d <= 2 && um->dp == 1;
  return (size_t)BB_STAGES * BB_TSTAGE * BB_TC * rec * 8 + 2 * BB_STAGES * 8 +
         (size_t)BB_WSTAGES * BB_THREADS * (BB_TC * um->dp) * 8 + (xbuf ? (size_t)BB_THREADS * 128 : 0);
}

int compile_cubin(bb_user_model* um, const kernel_key& k, std::vector<char>& cubin) {
  std::string err;
  int rc = rtc_load(err);
  if (rc) { um->log = err; return rc; }
  const std::string src = user_source(um, k);
  std::vector<const char*> hnames, htexts;
  for (int i = 0; i < bb_hdr_count; i++) { hnames.push_back(bb_hdr_names[i]); htexts.push_back(bb_hdr_texts[i]); }
  const char* empty_hdrs[] = {"cuda_runtime.h", "atomic", "math.h", "type_traits"};
  hnames.push_back("stdint.h"); htexts.push_back(k_stdint);
  for (const char* n : empty_hdrs) { hnames.push_back(n); htexts.push_back("\n"); }
  nvrtc_program prog = nullptr;
  int e = g_rtc.CreateProgram(&prog, src.c_str(), "bb_user_model.cu", (int)hnames.size(), htexts.data(), hnames.data());
  if (e) { um->log = std::string("nvrtcCreateProgram: ") + (g_rtc.GetErrorString ? g_rtc.GetErrorString(e) : "error"); return BB_ERR_UNSUPPORTED; }
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo", "-default-device"};
  e = g_rtc.CompileProgram(prog, (int)(sizeof(opts) / sizeof(opts[0])), opts);
  size_t ls = 0;
  if (g_rtc.GetProgramLogSize && g_rtc.GetProgramLogSize(prog, &ls) == 0 && ls > 1) {
    std::string lg(ls, '\0');
    g_rtc.GetProgramLog(prog, &lg[0]);
    um->log = lg;
  }
  if (e) {
    g_rtc.DestroyProgram(&prog);
    return BB_ERR_USERSRC;
  }
  size_t cs = 0;
  g_rtc.GetCUBINSize(prog, &cs);
  cubin.resize(cs);
  g_rtc.GetCUBIN(prog, cubin.data());
  g_rtc.DestroyProgram(&prog);
  return BB_OK;
}

int compile_kernel(bb_user_model* um, const kernel_key& k, jit_kernel& out) {
  std::vector<char> cubin;
  int rc = compile_cubin(um, k, cubin);
  if (rc) return rc;
  BB_CUDA(cudaLibraryLoadData(&out.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  BB_CUDA(cudaLibraryGetKernel(&out.kern, out.lib, "bb_user_kernel"));
  out.smem = k.kind == 0 ? path_smem(um, k) : 0;
  return BB_OK;
}
}  // namespace

bb_user_model* bb_user_lookup(int handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  return (handle > 0 && handle < (int)g_models.size()) ? g_models[handle] : nullptr;
}

/* launch the kernel of `um` for (kind, gk, gm, auxm, rng), compiling it on first use */
int bb_user_launch(bb_user_model* um, int kind, int gk, int gm, int auxm, int rng, const bb_chain_args& a, cudaStream_t st) {
  kernel_key k;
  memset(&k, 0, sizeof(k));
  k.kind = kind; k.gk = gk; k.gm = gm; k.auxm = auxm; k.rng = rng;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = um->kernels.find(k);
  if (it == um->kernels.end()) {
    jit_kernel jk;
    int rc = compile_kernel(um, k, jk);
    if (rc) return rc;
    it = um->kernels.emplace(k, jk).first;
  }
  jit_kernel& jk = it->second;
  if (!jk.attr_set && jk.smem > 48 * 1024) {
    BB_CUDA(cudaFuncSetAttribute((const void*)jk.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jk.smem));
    jk.attr_set = true;
  }
  const unsigned grid = (unsigned)((a.p_end - a.p_begin + BB_THREADS - 1) / BB_THREADS);
  bb_chain_args b = a;
  b.cpc = BB_THREADS; /* full CTAs */
  void* params[] = {&b};
  BB_CUDA(cudaLaunchKernel((const void*)jk.kern, dim3(grid), dim3(BB_THREADS), params, jk.smem, st));
  return BB_OK;
}

extern "C" int bb_user_model_create(bb_ctx* ctx, int32_t d, int32_t dprime, const char* drift_src, const int32_t* col,
                                    const char* const* sigma_src, bb_user_model** out) {
  if (!out) return BB_ERR_ARG;
  *out = nullptr;
  if (!ctx) return BB_ERR_NODEVICE;
  if (!drift_src || !col || d < 1 || d > 3 || dprime < 1 || dprime > d) return BB_ERR_ARG;
  for (int i = 0; i < d; i++) {
    if (col[i] < -1 || col[i] >= dprime) return BB_ERR_ARG;
    if (col[i] >= 0 && (!sigma_src || !sigma_src[i])) return BB_ERR_ARG;
  }
  bb_user_model* um = new (std::nothrow) bb_user_model();
  if (!um) return BB_ERR_NOMEM;
  um->ctx = ctx; um->d = d; um->dp = dprime; um->drift = drift_src;
  um->col.assign(col, col + d);
  um->sig.resize(d);
  for (int i = 0; i < d; i++)
    if (col[i] >= 0) um->sig[i] = sigma_src[i];
  {
    std::lock_guard<std::mutex> lk(g_mu);
    um->handle = (int)g_models.size();
    g_models.push_back(um);
  }
  ctx->live++;
  /* compile the plain Euler-Maruyama kernel now, so that errors in the user's source surface here with a log */
  bb_chain_args dummy;
  (void)dummy;
  kernel_key k;
  memset(&k, 0, sizeof(k));
  k.kind = 0; k.auxm = 1; k.rng = 0;
  jit_kernel jk;
  BB_CUDA(cudaSetDevice(ctx->device));
  int rc;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    rc = compile_kernel(um, k, jk);
    if (rc == BB_OK) um->kernels.emplace(k, jk);
  }
  *out = um; /* returned even on a compile error so that the caller can read the log; destroy it afterwards */
  return rc;
}
/* compile-only check of a user model's source (no device needed): the kernel for (guide kind, rows of L, auxiliary mode,
 * RNG mode) is built to a cubin and discarded; NVRTC's log goes to log[0..log_size) */
extern "C" int bb_user_source_check(int32_t d, int32_t dprime, const char* drift_src, const int32_t* col,
                                    const char* const* sigma_src, int32_t gk, int32_t gm, int32_t auxm, int32_t rng,
                                    char* log, int32_t log_size) {
  if (!drift_src || !col || d < 1 || d > 3 || dprime < 1 || dprime > d) return BB_ERR_ARG;
  bb_user_model um;
  um.d = d; um.dp = dprime; um.drift = drift_src;
  um.col.assign(col, col + d);
  um.sig.resize(d);
  for (int i = 0; i < d; i++) {
    if (col[i] < -1 || col[i] >= dprime) return BB_ERR_ARG;
    if (col[i] >= 0) {
      if (!sigma_src || !sigma_src[i]) return BB_ERR_ARG;
      um.sig[i] = sigma_src[i];
    }
  }
  kernel_key k;
  memset(&k, 0, sizeof(k));
  k.kind = rng >= 10 ? (rng == 12 ? 2 : 1) : 0;
  k.gk = gk; k.gm = gm; k.auxm = auxm; k.rng = rng >= 10 ? rng - 10 : rng;
  std::vector<char> cubin;
  int rc;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    rc = compile_cubin(&um, k, cubin);
  }
  if (log && log_size > 0) {
    strncpy(log, um.log.c_str(), (size_t)log_size - 1);
    log[log_size - 1] = '\0';
  }
  return rc == BB_OK ? (int)cubin.size() : rc;
}
extern "C" int32_t bb_user_model_handle(bb_user_model* um) { return um ? um->handle : 0; }
extern "C" const char* bb_user_model_log(bb_user_model* um) { return um ? um->log.c_str() : ""; }
extern "C" int bb_user_model_destroy(bb_user_model* um) {
  if (!um) return BB_ERR_ARG;
  cudaSetDevice(um->ctx->device);
  cudaStreamSynchronize(um->ctx->stream);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : um->kernels)
      if (kv.second.lib) cudaLibraryUnload(kv.second.lib);
    if (um->handle > 0 && um->handle < (int)g_models.size()) g_models[um->handle] = nullptr;
  }
  um->ctx->live--;
  delete um;
  return BB_OK;
}
