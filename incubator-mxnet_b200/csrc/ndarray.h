// ndarray.h -- the minimal NDArray the KVStore path needs: a ref-counted chunk of
// device (or pinned / pageable host) memory + shape/dtype/storage type, row-sparse
// aux indices, DLPack in/out.  Mirrors the subset of include/mxnet/ndarray.h:851-1122
// that src/kvstore touches; everything asynchronous is ordered on the owning
// device's engine stream (runtime.h).
#pragma once
#include <memory>
#include "base.h"
#include "runtime.h"

struct DLManagedTensor;

namespace mxkv {

struct Chunk {
  enum Kind { kOwnedCuda, kOwnedPinned, kOwnedHost, kSymmetric, kExternal, kDLPack, kExternalSymmetric };
  void* ptr = nullptr;
  size_t bytes = 0;
  Context ctx;
  Kind kind = kExternal;
  DLManagedTensor* dl = nullptr;
  SymPtr sym;                      // valid when kind == kSymmetric / kExternalSymmetric (MP mode)
  void* mc_ptr = nullptr;          // NVSwitch multicast alias of the same bytes on every rank, if bound
  ~Chunk();
};

class NDArray {
 public:
  NDArray() {}
  // dense allocation; symmetric=true allocates collectively from the process group's arena
  static NDArray Empty(const std::vector<int64_t>& shape, Context ctx, int dtype, bool symmetric = false);
  // row_sparse with capacity for `cap_rows` rows (indices int64)
  static NDArray EmptyRowSparse(const std::vector<int64_t>& shape, Context ctx, int dtype, int64_t cap_rows);
  static NDArray FromExternal(void* ptr, const std::vector<int64_t>& shape, Context ctx, int dtype);
  static NDArray FromDLPack(DLManagedTensor* dl, bool transient);
  DLManagedTensor* ToDLPack() const;

  bool is_none() const { return !chunk_; }
  void* data() const { return chunk_ ? static_cast<char*>(chunk_->ptr) + byte_offset_ : nullptr; }
  // pointer to the same bytes as mapped for peer `r` (symmetric chunks only)
  void* peer_data(int r) const;
  bool symmetric() const {
    return chunk_ && (chunk_->kind == Chunk::kSymmetric || chunk_->kind == Chunk::kExternalSymmetric);
  }
  // multicast (NVLS) address of element 0, or nullptr
  void* mc_data() const { return (chunk_ && chunk_->mc_ptr) ? static_cast<char*>(chunk_->mc_ptr) + byte_offset_ : nullptr; }
  // peer-mapped memory owned by the embedding framework (e.g. torch symmetric memory)
  static NDArray FromPeers(void* const* peer_ptrs, int world, int rank, void* mc_ptr,
                           const std::vector<int64_t>& shape, Context ctx, int dtype);
  const std::vector<int64_t>& shape() const { return shape_; }
  int64_t size() const { return ShapeSize(shape_); }
  size_t nbytes() const { return static_cast<size_t>(size()) * DTypeSize(dtype_); }
  int dtype() const { return dtype_; }
  int stype() const { return stype_; }
  Context ctx() const { return chunk_ ? chunk_->ctx : Context(); }
  int dev() const { return ctx().is_gpu() ? ctx().dev_id : -1; }

  NDArray Reshape(const std::vector<int64_t>& shape) const;
  NDArray Slice1D(int64_t begin, int64_t end) const;   // flat element range view

  // ---- row_sparse ----
  // logical shape_ is the full [num_rows, ...]; values live in a [cap_rows, row_len] chunk,
  // indices in aux_ (int64 [cap_rows]); nnz_ rows are initialised.  nnz_ < 0: the count
  // lives only on the device (d_nnz_) until someone asks (nnz() syncs).
  int64_t row_len() const;
  int64_t cap_rows() const { return cap_rows_; }
  // grow the row capacity (contents are discarded); row_sparse arrays are dynamically sized in the
  // reference (CheckAndAlloc), handles stay valid
  void ReserveRows(int64_t rows);
  int64_t nnz() const;
  void set_nnz(int64_t n) { *nnz_ = n; }
  void set_nnz_device() { *nnz_ = -1; }
  int64_t* d_nnz() const { return d_nnz_ ? static_cast<int64_t*>(d_nnz_->ptr) : nullptr; }
  NDArray aux_idx() const;         // dense int64 [nnz] view
  NDArray data_nd() const;         // dense [nnz, ...] view
  int64_t* idx_ptr() const { return aux_ ? static_cast<int64_t*>(aux_->ptr) : nullptr; }

  // synchronous host copies (MXNDArraySyncCopyFromCPU / ToCPU, c_api.h:778-800)
  void SyncCopyFromCPU(const void* src, size_t elems);
  void SyncCopyToCPU(void* dst, size_t elems) const;
  void WaitToRead() const;
  void WaitToWrite() const { WaitToRead(); }

  std::shared_ptr<Chunk> chunk_;
  size_t byte_offset_ = 0;
  std::vector<int64_t> shape_;
  int dtype_ = kFloat32;
  int stype_ = kDefaultStorage;
  std::shared_ptr<Chunk> aux_;
  std::shared_ptr<Chunk> d_nnz_;
  std::shared_ptr<int64_t> nnz_;
  int64_t cap_rows_ = 0;
};

// What an NDArrayHandle points to: the array plus per-handle scratch for the pointers
// MXNDArrayGetShape* hands out (the reference keeps those in MXAPIThreadLocalEntry).  Every
// handle that crosses the C ABI -- including the two the updater callback must free
// (c_api.cc:3066-3080) -- is allocated as an NDHandle and released by MXNDArrayFree.
struct NDHandle : public NDArray {
  std::vector<int> shape32;
  std::vector<int64_t> shape64;
  explicit NDHandle(const NDArray& a) : NDArray(a) {}
  NDHandle() {}
};

// async copy between any two dense arrays of equal byte size, ordered on engine streams
// (CopyFromTo, src/ndarray/ndarray.cc:1331-1424, without the per-op host wait)
void CopyFromTo(const NDArray& src, const NDArray& dst);
void CopyBytes(const void* src, Context sctx, void* dst, Context dctx, size_t bytes);

}  // namespace mxkv
