// ref_topology.cc -- compiles the REFERENCE's own tree solver (src/kvstore/gpu_topology.h, a header of
// templates over std::vector) into oracle/_ref/libkvref_topo.so, so that the engine's restated solver
// (incubator-mxnet_b200/csrc/topology.cc) can be pinned tree for tree against what libmxnet would build
// from the same link matrix.  TEST INFRASTRUCTURE ONLY.
//
// Nothing from the reference tree is copied: the header is #included where it lies (oracle/Makefile adds
// -I/root/reference and the dmlc-core include directory); this file supplies the three names the header
// expects its includer to have declared (mxnet::Context with a dev_id, dmlc::GetEnv, LOG) and C entry points
// over the functions the reference's own unit test calls (tests/cpp/kvstore/gpu_topology_test.cc).
// MXNET_USE_CUDA is 0: the solver itself is host code; only GetP2PWeight's device queries are compiled out.
// The build exists only in the authoring container; the built .so travels with the snapshot.
#include <dmlc/logging.h>
#include <dmlc/parameter.h>
#include <cstdint>
#include <cstring>

#define MXNET_USE_CUDA 0
namespace mxnet {
struct Context { int dev_id; };
}  // namespace mxnet
#include "src/kvstore/gpu_topology.h"

namespace ref = mxnet::kvstore;

extern "C" {

int kvref_topo_depth(int n) { return ref::ComputeDepth(n); }

// every tree flattened: tree i occupies topo_out[i * stride ...] with stride = 2^(depth+1) - 1 entries (the
// complete binary tree in array form), scan likewise with depth + 2 entries.  W is updated in place like
// CommDeviceTree's W_copy is not: pass a copy.  Returns 0, -1 when a tree has an unexpected size, -2 where the
// reference aborts (LOG(FATAL): no tree).
int kvref_topo_compute_trees(const float* W, int n, float alpha, int backtrack, uint64_t* topo_out,
                             uint64_t* scan_out, int* topo_len, int* scan_len) {
  std::vector<float> w(W, W + static_cast<size_t>(n) * n);
  std::vector<std::vector<size_t>> topo, scan;
  try {
    ref::ComputeTrees(w, n, alpha, backtrack != 0, &topo, &scan);
  } catch (const dmlc::Error&) {     // LOG(FATAL): no balanced binary tree over these links
    return -2;
  }
  size_t tl = topo.empty() ? 0 : topo[0].size(), sl = scan.empty() ? 0 : scan[0].size();
  for (int i = 0; i < n; ++i) {
    if (topo[i].size() != tl || scan[i].size() != sl) return -1;
    for (size_t j = 0; j < tl; ++j) topo_out[i * tl + j] = topo[i][j];
    for (size_t j = 0; j < sl; ++j) scan_out[i * sl + j] = scan[i][j];
  }
  *topo_len = static_cast<int>(tl);
  *scan_len = static_cast<int>(sl);
  return 0;
}

// one tree from one root on a matrix that is updated in place (the link-usage penalty), as the reference's
// TestComputeTreesFromRoot1 calls it
int kvref_topo_from_root(float* W, int n, int root, float alpha, int backtrack, uint64_t* topo_out, int* topo_len,
                         uint64_t* scan_out, int* scan_len) {
  std::vector<float> w(W, W + static_cast<size_t>(n) * n);
  std::vector<size_t> topo{static_cast<size_t>(root)}, scan{0};
  ref::ComputeTreesFromRoot(&w, n, root, alpha, backtrack != 0, &topo, &scan);
  for (size_t j = 0; j < topo.size(); ++j) topo_out[j] = topo[j];
  for (size_t j = 0; j < scan.size(); ++j) scan_out[j] = scan[j];
  *topo_len = static_cast<int>(topo.size());
  *scan_len = static_cast<int>(scan.size());
  std::memcpy(W, w.data(), w.size() * sizeof(float));
  return 0;
}

// one Kernighan-Lin pass over every cluster of P (TestKernighanLin1/2): returns `stop`
int kvref_topo_kernighan_lin(const float* W, int n, int* P, int* num_partitions, int* pairs_out, int* n_pairs,
                             uint32_t seed) {
  std::vector<float> w(W, W + static_cast<size_t>(n) * n);
  std::vector<int> p(P, P + n);
  std::vector<std::pair<int, int>> pairs;
  std::mt19937 gen(seed);
  const bool stop = ref::KernighanLin(w, &p, num_partitions, &pairs, &gen);
  for (int i = 0; i < n; ++i) P[i] = p[i];
  for (size_t i = 0; i < pairs.size(); ++i) { pairs_out[2 * i] = pairs[i].first; pairs_out[2 * i + 1] = pairs[i].second; }
  *n_pairs = static_cast<int>(pairs.size());
  return stop ? 1 : 0;
}

void kvref_topo_postprocess(int* result, int len, int n, int depth) {
  std::vector<int> r(result, result + len);
  ref::Postprocess(&r, n, depth);
  for (int i = 0; i < len; ++i) result[i] = r[i];
}

float kvref_topo_tree_weight(const float* W, const int* result, int len, int n, int depth, int penalty) {
  std::vector<float> w(W, W + static_cast<size_t>(n) * n);
  std::vector<int> r(result, result + len);
  return ref::ComputeTreeWeight(w, r, n, depth, penalty != 0);
}

int kvref_topo_is_valid(const float* W, const int* state, int len, int n, int row, int depth) {
  std::vector<float> w(W, W + static_cast<size_t>(n) * n);
  std::vector<int> s(state, state + len);
  return ref::IsValid(w, s, n, row, depth) ? 1 : 0;
}

int kvref_topo_is_connected(const float* W, int n) {
  std::vector<float> w(W, W + static_cast<size_t>(n) * n);
  return ref::IsConnected(w, n) ? 1 : 0;
}

}  // extern "C"
