// rsp_kernels.h -- launcher interface of the row_sparse kernels (rsp_kernels.cu).
#pragma once
#include "kernels.h"

namespace mxkv {

struct RspSources {
  const int64_t* idx[kMaxSrc];   // sorted unique row ids of source s
  const float* val[kMaxSrc];     // [nnz_s x row_len]
  const int64_t* nnz[kMaxSrc];   // device scalars
  int n;
};

struct RspRowArgs {
  int64_t* out_idx;        // [n * cap] sorted unique union of the ids
  float* out_val;          // [n * cap x row_len] merged values, or nullptr when not needed
  int64_t* d_nnz_out;      // device scalar: union size
  float* table;            // dense-backed stored value [num_rows x row_len] (optimizer / assign target)
  float* s0;               // momentum | adam mean, same layout as table
  float* s1;               // adam variance
  int64_t row_len;
  int opt;                 // OptKind (NONE, SGD, SGD_MOM, ADAM: lazy row-wise update)
  int assign;              // opt == NONE: also write the merged rows into the table
  int vec;                 // row_len % 4 == 0 and every pointer 16-byte aligned
  float lr, wd, rescale, clip, momentum, beta1, beta2, eps;
};

int LaunchRspSum(const RspSources& S, const RspRowArgs& A, int32_t* first, int32_t* pf, int64_t cap,
                 cudaStream_t stream);
int RspUniqueMax();
int LaunchRspUnique(const int64_t* ids, int64_t n, int64_t* out, int64_t* d_count, cudaStream_t stream);
int LaunchRspGather(const float* table, const int64_t* ids, const int64_t* d_count, int64_t max_rows, int64_t L,
                    float* out_val, int64_t* out_idx, int vec, cudaStream_t stream);
int LaunchRspScatter(float* table, const int64_t* idx, const int64_t* d_nnz, int64_t max_rows, int64_t L,
                     const float* val, cudaStream_t stream);
int LaunchSetI64(int64_t* p, int64_t v, cudaStream_t stream);
// dst[i] = (int64) src[i] for float32 / float64 / int32 / int64 row ids
int LaunchCastIdsToI64(const void* src, int dtype, int64_t* dst, int64_t n, cudaStream_t stream);
// cross-process rendezvous of block 0 (one-process-per-GPU mode, between non-collective kernels)
int LaunchBarrier(const SyncArgs& sync, cudaStream_t stream);

}  // namespace mxkv
