// ndarray.cc
#include "ndarray.h"
#include <cstdlib>
#include <cstring>
#include "dlpack_abi.h"

namespace mxkv {

Chunk::~Chunk() {
  // The reference defers NDArray deletion through the engine (Engine::DeleteVariable) so memory
  // outlives every pending reader.  Here kernels on ANY GPU may still be reading this buffer over
  // NVLink (cudaFree only synchronises the owning device), so drain the engine streams first.
  if ((kind == kOwnedCuda || kind == kOwnedPinned) && ptr) Runtime::Get()->DrainForFree();
  switch (kind) {
    case kOwnedCuda: {
      if (ptr) {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(ctx.dev_id);
        cudaFree(ptr);
        if (prev >= 0) cudaSetDevice(prev);
        cudaGetLastError();
      }
      break;
    }
    case kOwnedPinned: if (ptr) { cudaFreeHost(ptr); cudaGetLastError(); } break;
    case kOwnedHost: std::free(ptr); break;
    case kDLPack: if (dl && dl->deleter) dl->deleter(dl); break;
    case kSymmetric:   // bump-allocated from the arena; released with the process group
    case kExternalSymmetric:
    case kExternal: break;
  }
}

static std::shared_ptr<Chunk> AllocChunk(size_t bytes, Context ctx, bool symmetric) {
  auto c = std::make_shared<Chunk>();
  c->bytes = bytes;
  c->ctx = ctx;
  const size_t alloc = bytes == 0 ? 16 : bytes;
  if (ctx.dev_type == kGPU) {
    Runtime* rt = Runtime::Get();
    rt->Dev(ctx.dev_id);   // validates the device and fails loudly when there is no GPU
    if (symmetric && rt->pg() != nullptr) {
      MXKV_CHECK(rt->pg()->dev() == ctx.dev_id) << "symmetric arrays live on the process group's GPU";
      c->sym = rt->pg()->SymAlloc(alloc);
      c->ptr = c->sym.ptr[rt->pg()->rank()];
      c->mc_ptr = c->sym.mc;                 // engine-owned multicast alias (null on the cudaIpc fallback)
      c->kind = Chunk::kSymmetric;
    } else {
      DeviceGuard g(ctx.dev_id);
      CUDA_CALL(cudaMalloc(&c->ptr, alloc));
      c->kind = Chunk::kOwnedCuda;
    }
  } else if (ctx.dev_type == kCPUPinned) {
    MXKV_CHECK(Runtime::Get()->NumDevices() > 0) << "cpu_pinned memory needs a CUDA device";
    CUDA_CALL(cudaMallocHost(&c->ptr, alloc));
    c->kind = Chunk::kOwnedPinned;
  } else {
    c->ptr = std::malloc(alloc);
    MXKV_CHECK(c->ptr != nullptr) << "out of host memory";
    c->kind = Chunk::kOwnedHost;
  }
  return c;
}

NDArray NDArray::Empty(const std::vector<int64_t>& shape, Context ctx, int dtype, bool symmetric) {
  NDArray a;
  a.shape_ = shape;
  a.dtype_ = dtype;
  a.chunk_ = AllocChunk(static_cast<size_t>(ShapeSize(shape)) * DTypeSize(dtype), ctx, symmetric);
  return a;
}

NDArray NDArray::EmptyRowSparse(const std::vector<int64_t>& shape, Context ctx, int dtype, int64_t cap_rows) {
  MXKV_CHECK(shape.size() >= 1) << "row_sparse needs at least one dimension";
  NDArray a;
  a.shape_ = shape;
  a.dtype_ = dtype;
  a.stype_ = kRowSparseStorage;
  a.cap_rows_ = cap_rows;
  const int64_t L = a.row_len();
  a.chunk_ = AllocChunk(static_cast<size_t>(cap_rows * L) * DTypeSize(dtype), ctx, false);
  a.aux_ = AllocChunk(static_cast<size_t>(cap_rows) * sizeof(int64_t), ctx, false);
  a.d_nnz_ = AllocChunk(sizeof(int64_t), ctx, false);
  a.nnz_ = std::make_shared<int64_t>(0);
  return a;
}

NDArray NDArray::FromExternal(void* ptr, const std::vector<int64_t>& shape, Context ctx, int dtype) {
  NDArray a;
  a.shape_ = shape;
  a.dtype_ = dtype;
  a.chunk_ = std::make_shared<Chunk>();
  a.chunk_->ptr = ptr;
  a.chunk_->bytes = static_cast<size_t>(ShapeSize(shape)) * DTypeSize(dtype);
  a.chunk_->ctx = ctx;
  a.chunk_->kind = Chunk::kExternal;
  return a;
}

NDArray NDArray::FromPeers(void* const* peer_ptrs, int world, int rank, void* mc_ptr,
                           const std::vector<int64_t>& shape, Context ctx, int dtype) {
  MXKV_CHECK(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world) << "bad peer table";
  NDArray a = FromExternal(peer_ptrs[rank], shape, ctx, dtype);
  a.chunk_->kind = Chunk::kExternalSymmetric;
  for (int r = 0; r < world; ++r) a.chunk_->sym.ptr[r] = peer_ptrs[r];
  a.chunk_->sym.valid = true;
  a.chunk_->mc_ptr = mc_ptr;
  return a;
}

static int DLToDType(const DLDataType& t) {
  MXKV_CHECK(t.lanes == 1) << "vector DLPack types are not supported";
  if (t.code == kDLFloat) {
    if (t.bits == 32) return kFloat32;
    if (t.bits == 64) return kFloat64;
    if (t.bits == 16) return kFloat16;
  } else if (t.code == kDLBfloat && t.bits == 16) {
    return kBfloat16;
  } else if (t.code == kDLInt) {
    if (t.bits == 32) return kInt32;
    if (t.bits == 64) return kInt64;
    if (t.bits == 8) return kInt8;
  } else if (t.code == kDLUInt && t.bits == 8) {
    return kUint8;
  }
  MXKV_FATAL() << "unsupported DLPack dtype code=" << int(t.code) << " bits=" << int(t.bits);
}

static DLDataType DTypeToDL(int dtype) {
  DLDataType t; t.lanes = 1;
  switch (dtype) {
    case kFloat32: t.code = kDLFloat; t.bits = 32; break;
    case kFloat64: t.code = kDLFloat; t.bits = 64; break;
    case kFloat16: t.code = kDLFloat; t.bits = 16; break;
    case kBfloat16: t.code = kDLBfloat; t.bits = 16; break;
    case kInt32: t.code = kDLInt; t.bits = 32; break;
    case kInt64: t.code = kDLInt; t.bits = 64; break;
    case kInt8: t.code = kDLInt; t.bits = 8; break;
    case kUint8: t.code = kDLUInt; t.bits = 8; break;
    default: MXKV_FATAL() << "dtype " << dtype << " has no DLPack encoding";
  }
  return t;
}

NDArray NDArray::FromDLPack(DLManagedTensor* dl, bool transient) {
  MXKV_CHECK(dl != nullptr) << "null DLManagedTensor";
  const DLTensor& t = dl->dl_tensor;
  NDArray a;
  a.shape_.assign(t.shape, t.shape + t.ndim);
  if (t.strides != nullptr) {   // must be compact row-major (CopyFromTo needs contiguity, ndarray_function.cu:86-89)
    int64_t expect = 1;
    for (int i = t.ndim - 1; i >= 0; --i) {
      MXKV_CHECK(t.shape[i] == 1 || t.strides[i] == expect) << "DLPack tensor is not contiguous";
      expect *= t.shape[i];
    }
  }
  a.dtype_ = DLToDType(t.dtype);
  a.chunk_ = std::make_shared<Chunk>();
  a.chunk_->ptr = t.data;
  a.byte_offset_ = static_cast<size_t>(t.byte_offset);
  a.chunk_->bytes = a.byte_offset_ + a.nbytes();
  if (t.device.device_type == kDLCUDA) a.chunk_->ctx = Context{kGPU, t.device.device_id};
  else if (t.device.device_type == kDLCUDAHost) a.chunk_->ctx = Context{kCPUPinned, 0};
  else if (t.device.device_type == kDLCPU) a.chunk_->ctx = Context{kCPU, 0};
  else MXKV_FATAL() << "unsupported DLPack device type " << int(t.device.device_type);
  if (transient) {
    a.chunk_->kind = Chunk::kExternal;
  } else {
    a.chunk_->kind = Chunk::kDLPack;
    a.chunk_->dl = dl;
  }
  return a;
}

namespace {
struct DLHolder {
  NDArray keep;
  DLManagedTensor t;
  std::vector<int64_t> shape;
};
void DLHolderDeleter(DLManagedTensor* t) { delete static_cast<DLHolder*>(t->manager_ctx); }
}  // namespace

DLManagedTensor* NDArray::ToDLPack() const {
  MXKV_CHECK(stype_ == kDefaultStorage) << "only dense arrays export to DLPack";
  DLHolder* h = new DLHolder();
  h->keep = *this;
  h->shape = shape_;
  h->t.manager_ctx = h;
  h->t.deleter = DLHolderDeleter;
  DLTensor& d = h->t.dl_tensor;
  d.data = chunk_->ptr;
  d.byte_offset = byte_offset_;
  d.ndim = static_cast<int>(shape_.size());
  d.shape = h->shape.data();
  d.strides = nullptr;
  d.dtype = DTypeToDL(dtype_);
  const Context c = ctx();
  d.device.device_type = c.dev_type == kGPU ? kDLCUDA : (c.dev_type == kCPUPinned ? kDLCUDAHost : kDLCPU);
  d.device.device_id = c.dev_type == kGPU ? c.dev_id : 0;
  return &h->t;
}

void* NDArray::peer_data(int r) const {
  MXKV_CHECK(symmetric() && chunk_->sym.ptr[r] != nullptr) << "array is not symmetric";
  return static_cast<char*>(chunk_->sym.ptr[r]) + byte_offset_;
}

NDArray NDArray::Reshape(const std::vector<int64_t>& shape) const {
  MXKV_CHECK(ShapeSize(shape) == size()) << "reshape changes the number of elements";
  NDArray a = *this;
  a.shape_ = shape;
  return a;
}

NDArray NDArray::Slice1D(int64_t begin, int64_t end) const {
  MXKV_CHECK(stype_ == kDefaultStorage && begin >= 0 && begin <= end && end <= size()) << "bad slice";
  NDArray a = *this;
  a.byte_offset_ += static_cast<size_t>(begin) * DTypeSize(dtype_);
  a.shape_ = {end - begin};
  return a;
}

void NDArray::ReserveRows(int64_t rows) {
  MXKV_CHECK(stype_ == kRowSparseStorage) << "ReserveRows on a dense array";
  if (rows <= cap_rows_) return;
  NDArray fresh = EmptyRowSparse(shape_, ctx(), dtype_, rows);
  chunk_ = fresh.chunk_;
  aux_ = fresh.aux_;
  byte_offset_ = 0;
  cap_rows_ = rows;
  *nnz_ = 0;
}

int64_t NDArray::row_len() const {
  int64_t L = 1;
  for (size_t i = 1; i < shape_.size(); ++i) L *= shape_[i];
  return L;
}

int64_t NDArray::nnz() const {
  MXKV_CHECK(stype_ == kRowSparseStorage) << "nnz() on a dense array";
  if (*nnz_ < 0) {   // count still on the device: one deliberate sync, on demand only
    int64_t v = 0;
    const Context c = ctx();
    if (c.is_gpu()) {
      DeviceGuard g(c.dev_id);
      DeviceState& d = Runtime::Get()->Dev(c.dev_id);
      CUDA_CALL(cudaMemcpyAsync(&v, d_nnz_->ptr, sizeof(v), cudaMemcpyDeviceToHost, d.stream));
      CUDA_CALL(cudaStreamSynchronize(d.stream));
    } else {
      Runtime::Get()->WaitAll();
      v = *static_cast<int64_t*>(d_nnz_->ptr);
    }
    *nnz_ = v;
  }
  return *nnz_;
}

NDArray NDArray::aux_idx() const {
  NDArray a;
  a.chunk_ = aux_;
  a.shape_ = {nnz()};
  a.dtype_ = kInt64;
  return a;
}

NDArray NDArray::data_nd() const {
  NDArray a;
  a.chunk_ = chunk_;
  a.byte_offset_ = byte_offset_;
  a.shape_ = shape_;
  a.shape_[0] = nnz();
  a.dtype_ = dtype_;
  return a;
}

void CopyBytes(const void* src, Context sctx, void* dst, Context dctx, size_t bytes) {
  if (bytes == 0 || src == dst) return;
  Runtime* rt = Runtime::Get();
  if (sctx.is_gpu() && dctx.is_gpu()) {
    if (sctx.dev_id == dctx.dev_id) {
      DeviceGuard g(sctx.dev_id);
      CUDA_CALL(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, rt->Dev(sctx.dev_id).stream));
    } else {
      rt->StreamWait(sctx.dev_id, dctx.dev_id);   // earlier readers/writers of dst
      {
        DeviceGuard g(sctx.dev_id);
        CUDA_CALL(cudaMemcpyPeerAsync(dst, dctx.dev_id, src, sctx.dev_id, bytes, rt->Dev(sctx.dev_id).stream));
      }
      rt->StreamWait(dctx.dev_id, sctx.dev_id);
    }
  } else if (sctx.is_gpu()) {
    DeviceGuard g(sctx.dev_id);
    CUDA_CALL(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, rt->Dev(sctx.dev_id).stream));
  } else if (dctx.is_gpu()) {
    DeviceGuard g(dctx.dev_id);
    CUDA_CALL(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, rt->Dev(dctx.dev_id).stream));
  } else {
    rt->WaitAll();   // pending async copies into src must have landed
    std::memcpy(dst, src, bytes);
  }
}

void CopyFromTo(const NDArray& src, const NDArray& dst) {
  MXKV_CHECK(src.stype() == kDefaultStorage && dst.stype() == kDefaultStorage) << "dense CopyFromTo only";
  MXKV_CHECK(src.size() == dst.size()) << "CopyFromTo: size mismatch " << src.size() << " vs " << dst.size();
  MXKV_CHECK(src.dtype() == dst.dtype()) << "CopyFromTo: dtype mismatch";
  CopyBytes(src.data(), src.ctx(), dst.data(), dst.ctx(), src.nbytes());
}

void NDArray::SyncCopyFromCPU(const void* src, size_t elems) {
  MXKV_CHECK(stype_ == kDefaultStorage) << "SyncCopyFromCPU: dense only";
  MXKV_CHECK(static_cast<int64_t>(elems) == size()) << "SyncCopyFromCPU: size mismatch";
  const Context c = ctx();
  if (c.is_gpu()) {
    DeviceGuard g(c.dev_id);
    DeviceState& d = Runtime::Get()->Dev(c.dev_id);
    CUDA_CALL(cudaMemcpyAsync(data(), src, nbytes(), cudaMemcpyHostToDevice, d.stream));
    CUDA_CALL(cudaStreamSynchronize(d.stream));
  } else {
    Runtime::Get()->WaitAll();
    std::memcpy(data(), src, nbytes());
  }
}

void NDArray::SyncCopyToCPU(void* dst, size_t elems) const {
  MXKV_CHECK(stype_ == kDefaultStorage) << "SyncCopyToCPU: dense only";
  MXKV_CHECK(static_cast<int64_t>(elems) == size()) << "SyncCopyToCPU: size mismatch";
  const Context c = ctx();
  if (c.is_gpu()) {
    DeviceGuard g(c.dev_id);
    DeviceState& d = Runtime::Get()->Dev(c.dev_id);
    CUDA_CALL(cudaMemcpyAsync(dst, data(), nbytes(), cudaMemcpyDeviceToHost, d.stream));
    CUDA_CALL(cudaStreamSynchronize(d.stream));
  } else {
    Runtime::Get()->WaitAll();
    std::memcpy(dst, data(), nbytes());
  }
}

void NDArray::WaitToRead() const {
  const Context c = ctx();
  if (c.is_gpu()) Runtime::Get()->WaitDevice(c.dev_id);
  else Runtime::Get()->WaitAll();
}

}  // namespace mxkv
