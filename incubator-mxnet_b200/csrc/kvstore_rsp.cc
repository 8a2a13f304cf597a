// kvstore_rsp.cc -- row_sparse push / row_sparse_pull (placeholder until the sparse kernels land).
#include "kvstore.h"

namespace mxkv {

void KVStore::InitRowSparseKey(KeyState& ks, const NDArray& v) {
  (void)ks; (void)v;
  MXKV_FATAL() << "row_sparse keys are not implemented yet";
}
void KVStore::PushRowSparse(KeyState& ks, const std::vector<NDArray>& vals) {
  (void)ks; (void)vals;
  MXKV_FATAL() << "row_sparse push is not implemented yet";
}
void KVStore::PullDenseFromRowSparse(KeyState& ks, const std::vector<NDArray*>& outs) {
  (void)ks; (void)outs;
  MXKV_FATAL() << "pull from a row_sparse key is not implemented yet";
}
void KVStore::PullRowSparseImpl(const std::vector<int>& keys,
                                const std::vector<std::pair<NDArray*, NDArray>>& vr, int priority) {
  (void)keys; (void)vr; (void)priority;
  MXKV_FATAL() << "row_sparse_pull is not implemented yet";
}

}  // namespace mxkv
