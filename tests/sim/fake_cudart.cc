// fake_cudart.cc -- a stand-in for libcudart.so.12 that lets the engine's HOST code run on a machine
// without a GPU (test infrastructure only; see tests/sim/README.md).
//
// "Device" memory is host memory, every "GPU" lives in this process and can reach every other one,
// streams execute synchronously in issue order, events are no-ops.  A kernel launch is dispatched by the
// kernel's (demangled) name to a semantic emulator in sim_kernels.cc, which performs what the kernel is
// specified to do with the launch descriptor it was given -- using the very arithmetic headers the real
// kernels are compiled from (csrc/optim_math.h, csrc/norm_math.h), so results are bit-identical to the
// oracle.  What this exercises is everything AROUND the kernels: key bookkeeping, placement, replica and
// state management, work-list construction, aliasing rules, launch sequencing, the C ABI and the Python
// front-end.  It cannot find a bug inside a CUDA kernel; the `-m gpu` tests on real hardware do that.
//
// The number of simulated GPUs is MXKV_SIM_DEVICES (default 4).
#include <cuda_runtime_api.h>
#include <cxxabi.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include "sim.h"

namespace {

std::mutex g_mu;
std::map<const void*, std::string> g_kernels;     // host stub address -> demangled kernel name
thread_local int g_device = 0;
thread_local cudaError_t g_last = cudaSuccess;

struct CallConfig { dim3 grid, block; size_t smem; void* stream; };
thread_local CallConfig g_cfg;

int device_count() {
  const char* e = getenv("MXKV_SIM_DEVICES");
  const int n = (e && *e) ? atoi(e) : 4;
  return n < 1 ? 1 : (n > 8 ? 8 : n);
}

cudaError_t fail(cudaError_t e) { g_last = e; return e; }

void* sim_alloc(size_t bytes) {
  void* p = nullptr;
  if (posix_memalign(&p, 512, bytes ? bytes : 16) != 0) return nullptr;
  return p;
}

}  // namespace

#define API extern "C" __attribute__((visibility("default")))

// ---- registration of the kernels compiled into the engine -----------------------------------
API void** __cudaRegisterFatBinary(void*) { static void* h = nullptr; return &h; }
API void __cudaRegisterFatBinaryEnd(void**) {}
API void __cudaUnregisterFatBinary(void**) {}
API void __cudaRegisterFunction(void**, const char* hostFun, char*, const char* deviceName, int, uint3*, uint3*,
                                dim3*, dim3*, int*) {
  int status = 0;
  char* dem = abi::__cxa_demangle(deviceName, nullptr, nullptr, &status);
  std::lock_guard<std::mutex> lk(g_mu);
  g_kernels[hostFun] = (status == 0 && dem) ? dem : deviceName;
  free(dem);
}
API void __cudaRegisterVar(void**, char*, char*, const char*, int, size_t, int, int) {}
API unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t smem, void* stream) {
  g_cfg = CallConfig{grid, block, smem, stream};
  return 0;
}
API cudaError_t __cudaPopCallConfiguration(dim3* grid, dim3* block, size_t* smem, void* stream) {
  *grid = g_cfg.grid; *block = g_cfg.block; *smem = g_cfg.smem;
  *static_cast<void**>(stream) = g_cfg.stream;
  return cudaSuccess;
}

API cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t) {
  std::string name;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_kernels.find(func);
    if (it == g_kernels.end()) return fail(cudaErrorInvalidDeviceFunction);
    name = it->second;
  }
  static const bool noexec = getenv("MXKV_SIM_NOEXEC") != nullptr;   // host-overhead measurements: launch = no-op
  if (noexec) return cudaSuccess;
  sim::LaunchInfo info;
  info.name = name;
  info.grid = grid.x; info.block = block.x; info.smem = smem;
  info.device = g_device;
  if (!sim::Dispatch(info, args)) {
    fprintf(stderr, "[mxkv sim] no emulator for kernel: %s\n", name.c_str());
    return fail(cudaErrorNotSupported);
  }
  return cudaSuccess;
}

// ---- devices ------------------------------------------------------------------------------------
API cudaError_t cudaGetDeviceCount(int* n) { *n = device_count(); return cudaSuccess; }
API cudaError_t cudaGetDevice(int* d) { *d = g_device; return cudaSuccess; }
API cudaError_t cudaSetDevice(int d) {
  if (d < 0 || d >= device_count()) return fail(cudaErrorInvalidDevice);
  g_device = d;
  return cudaSuccess;
}
API cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
API cudaError_t cudaDeviceCanAccessPeer(int* can, int a, int b) { *can = (a != b) ? 1 : 0; return cudaSuccess; }
API cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
API cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -5; return cudaSuccess; }
API cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr attr, int) {
  switch (attr) {
    case cudaDevAttrMultiProcessorCount: *v = 148; break;
    case cudaDevAttrClockRate: *v = 1965000; break;
    case cudaDevAttrMaxSharedMemoryPerBlockOptin: *v = 227 * 1024; break;
    default: *v = 0; break;
  }
  return cudaSuccess;
}
API cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* pool, int) { *pool = nullptr; return cudaSuccess; }
API cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }
API cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
API cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int* n, const void*, int, size_t, unsigned) {
  *n = 2;
  return cudaSuccess;
}

// ---- errors ---------------------------------------------------------------------------------------
API cudaError_t cudaGetLastError(void) { const cudaError_t e = g_last; g_last = cudaSuccess; return e; }
API const char* cudaGetErrorString(cudaError_t e) {
  switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorNotSupported: return "operation not supported (simulated runtime)";
    case cudaErrorInvalidValue: return "invalid argument";
    case cudaErrorInvalidDevice: return "invalid device ordinal";
    case cudaErrorInvalidDeviceFunction: return "invalid device function";
    default: return "simulated CUDA error";
  }
}

// ---- memory -----------------------------------------------------------------------------------------
API cudaError_t cudaMalloc(void** p, size_t n) { *p = sim_alloc(n); return *p ? cudaSuccess : fail(cudaErrorMemoryAllocation); }
API cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
API cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
API cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
API cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
API cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
API cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
API cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t) {
  if (n) memmove(dst, src, n);
  return cudaSuccess;
}
API cudaError_t cudaMemcpyPeerAsync(void* dst, int, const void* src, int, size_t n, cudaStream_t) {
  if (n) memmove(dst, src, n);
  return cudaSuccess;
}
API cudaError_t cudaMemset(void* p, int v, size_t n) { if (n) memset(p, v, n); return cudaSuccess; }
API cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { return cudaMemset(p, v, n); }
API cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return fail(cudaErrorNotSupported); }
API cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return fail(cudaErrorNotSupported); }
API cudaError_t cudaIpcCloseMemHandle(void*) { return fail(cudaErrorNotSupported); }

// ---- streams and events: everything has already happened -----------------------------------------
API cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  *s = reinterpret_cast<cudaStream_t>(sim_alloc(16));
  return cudaSuccess;
}
API cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned f, int) { return cudaStreamCreateWithFlags(s, f); }
API cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
API cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
API cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
  *e = reinterpret_cast<cudaEvent_t>(sim_alloc(16));
  return cudaSuccess;
}
API cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
API cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
API cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
API cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
