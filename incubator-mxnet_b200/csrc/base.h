// base.h -- error plumbing and small shared types of the host runtime.
//
// Error contract mirrors the reference C API: every entry point returns 0 / -1 and
// never throws across the boundary; the message is kept per thread and read with
// MXGetLastError() (include/mxnet/c_api_error.h:40-58, src/runtime/c_runtime_api.cc:253-264).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "kernels.h"

namespace mxkv {

struct Error : public std::runtime_error {
  explicit Error(const std::string& s) : std::runtime_error(s) {}
};

void SetLastError(const std::string& msg);
const char* GetLastError();

struct ErrStream {
  std::ostringstream os;
  template <typename T> ErrStream& operator<<(const T& v) { os << v; return *this; }
};
struct ErrThrow {
  // `ErrThrow() & stream` throws after the message has been assembled
  [[noreturn]] void operator&(const ErrStream& s) { throw Error(s.os.str()); }
};

#define MXKV_CHECK(cond)                                                         \
  if (cond) {} else ::mxkv::ErrThrow() & ::mxkv::ErrStream()                      \
      << "Check failed: " #cond " (" << __FILE__ << ":" << __LINE__ << ") "

#define MXKV_FATAL() ::mxkv::ErrThrow() & ::mxkv::ErrStream()

#define CUDA_CALL(expr)                                                          \
  do {                                                                           \
    cudaError_t e__ = (expr);                                                    \
    if (e__ != cudaSuccess && e__ != cudaErrorCudartUnloading) {                  \
      cudaGetLastError();                                                        \
      MXKV_FATAL() << "CUDA: " << cudaGetErrorString(e__) << " at " << __FILE__  \
                   << ":" << __LINE__ << " in " #expr;                           \
    }                                                                            \
  } while (0)

// Context::DeviceType, include/mxnet/base.h:94-99
enum DevType : int { kCPU = 1, kGPU = 2, kCPUPinned = 3 };
// NDArrayStorageType, include/mxnet/ndarray.h:61-66
enum StorageType : int { kUndefinedStorage = -1, kDefaultStorage = 0, kRowSparseStorage = 1, kCSRStorage = 2 };

struct Context {
  int dev_type = kCPU;
  int dev_id = 0;
  bool is_gpu() const { return dev_type == kGPU; }
  bool operator==(const Context& o) const { return dev_type == o.dev_type && dev_id == o.dev_id; }
  bool operator!=(const Context& o) const { return !(*this == o); }
};

inline size_t DTypeSize(int dtype) {
  switch (dtype) {
    case kFloat32: case kInt32: case kUint32: return 4;
    case kFloat64: case kInt64: case kUint64: return 8;
    case kFloat16: case kBfloat16: case kInt16: case kUint16: return 2;
    case kUint8: case kInt8: case kBool: return 1;
  }
  MXKV_FATAL() << "Unknown type enum " << dtype;
}

inline int64_t ShapeSize(const std::vector<int64_t>& s) {
  int64_t n = 1;
  for (auto d : s) n *= d;
  return n;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (dev >= 0 && dev != prev) CUDA_CALL(cudaSetDevice(dev));
    else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int64_t EnvInt(const char* name, int64_t dflt);

}  // namespace mxkv
