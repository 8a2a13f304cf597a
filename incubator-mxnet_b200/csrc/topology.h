// topology.h -- the reduction trees of MXNET_KVSTORE_USETREE=1 (reference: CommDeviceTree,
// src/kvstore/comm_tree.h:50-325, and its solver src/kvstore/gpu_topology.h:137-1157).
//
// What the reference's tree mode changes that a caller can observe is the ASSOCIATION of the gradient sum: the
// n replicas of a key are added pairwise up a binary tree whose shape comes from the machine's link matrix, and
// a key above MXNET_KVSTORE_TREE_ARRAY_BOUND elements is cut into n row slices with slice i summed up the tree
// rooted at GPU i (comm_tree.h:203-234).  Everything else -- which GPU holds which partial, the log2(n) copy +
// sum passes -- is transport.  So this file
//   1. derives the link matrix the reference would see (LinkWeights: GetP2PWeight, gpu_topology.h:137-253),
//   2. builds the n trees from it, tree for tree what the reference builds (ComputeTrees: Kernighan-Lin
//      bisection + edge picking, gpu_topology.h:326-700, or the exhaustive search, :846-1016; same random
//      engine, seed and draw order, so the same trees on the same C++ library),
//   3. turns one tree into a *reduce program* (ReduceProgramOf): the order in which one kernel thread must add
//      the n replicas of an element to produce the bits CommDeviceTree::ReduceInner (comm_tree.h:91-177) would,
// and the dense kernel (kernels.cu, TREE instantiations) executes that program per element while it streams the
// replicas over NVLink -- one pass, no merge buffers, transport unchanged.
#pragma once
#include <stdint.h>
#include <utility>
#include <vector>
#include "kernels.h"

namespace mxkv {
namespace topo {

// Trees in the reference's layout: tree r (rooted at participant r) is a complete binary tree of `depth` levels
// below the root stored level by level -- node i has children 2i+1, 2i+2 and a parent is repeated as its own
// left child; a pair (g, g) is a GPU that sits a level out -- and scan[r][l] is where level l starts.
struct TreeSet {
  int n = 0;
  int depth = 0;
  std::vector<std::vector<size_t>> topo;
  std::vector<std::vector<size_t>> scan;
};

// smallest number of levels d >= 1 with n <= 2^d (ComputeDepth, gpu_topology.h:714-721); 0 above 2^16
int TreeDepth(int n);

// perf_rank[r*n+c]: cudaDeviceGetP2PAttribute(cudaDevP2PAttrPerformanceRank, dev r, dev c); can_access: 1 where peer
// access could be enabled.  Returns the n x n weight matrix of GetP2PWeight (gpu_topology.h:137-253).
std::vector<float> LinkWeights(int n, const std::vector<int>& perf_rank, const std::vector<int>& can_access);
bool LinksConnected(const std::vector<float>& W, int n);      // IsConnected, gpu_topology.h:96-121

// ComputeTrees, gpu_topology.h:1111-1157.  Throws mxkv::Error where the reference aborts (no balanced binary tree).
void ComputeTrees(const std::vector<float>& W, int n, float alpha, bool backtrack, TreeSet* out);

// pieces of the solver the reference's unit test pins one by one (tests/cpp/kvstore/gpu_topology_test.cc)
bool BisectClusters(const std::vector<float>& W, std::vector<int>* color, int* n_colors,
                    std::vector<std::pair<int, int>>* pairs, void* mt19937_engine);
void FoldRepeats(std::vector<int>* leaves, int n, int depth);                              // Postprocess, :746-770
float TreeWeight(const std::vector<float>& W, const std::vector<int>& leaves, int n, int depth, bool penalty);
bool Admissible(const std::vector<float>& W, const std::vector<int>& state, int n, int row, int depth);

// The sum CommDeviceTree::ReduceInner computes up one tree, as a straight-line program for one thread:
// visit the participants in `leaf` order; after taking participant leaf[k]'s value, read the low bits of `prog`:
// every 0 bit adds the newest pending partial sum to the running value (one ElementwiseSum of the reference),
// the first 1 bit ends the step and the running value becomes the newest pending partial.  n values need n 1-bits
// and n - 1 0-bits; at most kTreeStack partials are pending at any time (trees of depth <= 3, i.e. n <= 8).
struct ReduceProgram {
  int n = 0;
  int leaf[kMaxRanks] = {0};
  uint32_t prog = 0;
};
ReduceProgram ReduceProgramOf(const std::vector<size_t>& topo_row, const std::vector<size_t>& scan_row, int depth, int n);

// devs: CUDA device ordinal of participant 0..n-1.  Queries the driver like CommDeviceTree::QueryTopology
// (comm_tree.h:370-385); MXKV_B200_TREE_LINKS ("w00,w01,...", n*n numbers) replaces the answer (tests).
std::vector<float> QueryLinkWeights(const std::vector<int>& devs);

}  // namespace topo
}  // namespace mxkv
