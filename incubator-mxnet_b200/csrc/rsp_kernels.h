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

// fused push (rsp_push_fused_kernel): staging of this rank's gradient and local copies of the id lists
struct RspStage {
  const int64_t* src_idx;  // this rank's gradient as the caller holds it ...
  const float* src_val;
  int64_t src_nnz;
  int64_t* dst_idx;        // ... and its peer-mapped staging area (one process per GPU)
  float* dst_val;
  int64_t* dst_nnz;
  int64_t* lidx;           // [n x lcap] local copies of every source's id list
  int64_t lcap;
  int64_t nnz_val[kMaxSrc]; // row counts by value (single process: the host knows them; saves a launch per value)
  int nnz_by_value;
  int publish;             // 1: phase P0 / P3 (copy into the staging area, cross-GPU barriers)
  int localize;            // 1: phase P1 (some source lives on another GPU)
};

// One launch: (publish) | (localize ids) | union + gather-sum + lazy update of the touched rows | (barrier).
// Only for pushes that do not have to materialise the merged value (A.out_val == nullptr, A.assign == 0).
// `est_rows`: the host's estimate of the candidate rows (sum of the sources' row counts), sizes the grid.
int LaunchRspPushFused(int device, const RspSources& S, const RspRowArgs& A, const RspStage& St, const SyncArgs& sync,
                       int64_t est_rows, cudaStream_t stream);
// One launch: unique (n <= RspUniqueMax()) | gather
int LaunchRspPullFused(int device, const float* table, const int64_t* ids, int64_t n, int64_t* out_idx, int64_t* d_count,
                       int64_t L, float* out_val, int vec, const SyncArgs& sync, cudaStream_t stream);

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
