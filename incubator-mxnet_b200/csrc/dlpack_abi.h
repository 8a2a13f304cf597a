// dlpack_abi.h -- the (stable, public) DLPack C ABI structs, declared locally so the
// library has no header dependency.  Layout per the DLPack specification v0.x
// (the unversioned DLManagedTensor that torch.utils.dlpack and
// MXNDArrayToDLPack/FromDLPack, include/mxnet/c_api.h:976-1002, exchange).
#pragma once
#include <stdint.h>

extern "C" {
typedef enum {
  kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3, kDLOpenCL = 4, kDLVulkan = 7, kDLMetal = 8,
  kDLVPI = 9, kDLROCM = 10, kDLROCMHost = 11, kDLExtDev = 12, kDLCUDAManaged = 13
} DLDeviceType;

typedef struct { int32_t device_type; int32_t device_id; } DLDevice;

typedef enum { kDLInt = 0U, kDLUInt = 1U, kDLFloat = 2U, kDLOpaqueHandle = 3U, kDLBfloat = 4U,
               kDLComplex = 5U, kDLBool = 6U } DLDataTypeCode;

typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;

typedef struct {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides;
  uint64_t byte_offset;
} DLTensor;

typedef struct DLManagedTensor {
  DLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;
}
