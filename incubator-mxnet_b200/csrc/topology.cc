// topology.cc -- see topology.h.  Host code only (compiled by nvcc with the rest of the engine, and by g++ for
// the CPU test of the solver: tests/c/topology_host.cc).
//
// The solver has to land on the SAME trees as the reference's, because the trees decide how sums are rounded.
// That pins more than the algorithm: the random engine (std::mt19937 seeded with 1 per root), every
// std::shuffle call and its position in the draw sequence, float (not double) gains, strict comparisons that
// make the first of several equal candidates win, and two carry-overs between refinement rounds of one cluster
// (the best gain and its prefix length are not reset) all change which of several equally good trees comes
// out.  Each is marked below with the reference lines it follows; tests/test_topology.py holds the solver to
// the reference's compiled header (oracle/_ref/libkvref_topo.so) and to committed trees (tests/golden).
#include "topology.h"
#include <algorithm>
#include <cstdlib>
#include <limits>
#include <queue>
#include <random>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include "base.h"

namespace mxkv {
namespace topo {

int TreeDepth(int n) {
  for (int d = 1; d <= 16; ++d)
    if (n <= (1 << d)) return d;
  return 0;
}

bool LinksConnected(const std::vector<float>& W, int n) {
  // breadth-first over the links that are better than PCI-E (weight > 1), from participant 0
  std::vector<char> seen(n, 0);
  std::queue<int> todo;
  todo.push(0);
  seen[0] = 1;
  while (!todo.empty()) {
    const int at = todo.front();
    todo.pop();
    for (int i = 0; i < n; ++i) {
      const int w = static_cast<int>(W[at * n + i]);          // the reference compares the truncated weight
      if (i != at && w > 1 && !seen[i]) { seen[i] = 1; todo.push(i); }
    }
  }
  for (int i = 0; i < n; ++i) if (!seen[i]) return false;
  return true;
}

std::vector<float> LinkWeights(int n, const std::vector<int>& perf_rank, const std::vector<int>& can_access) {
  std::vector<float> W(static_cast<size_t>(n) * n, 0.f);
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c)
      W[r * n + c] = r == c ? 0.f : static_cast<float>(perf_rank[r * n + c]) + 1.f;
  // a pair that can map each other's memory but ranks like PCI-E: the attribute is not to be trusted
  // (gpu_topology.h:173-221); fall back to "peer access = one NVLink"
  bool trusted = true;
  for (size_t i = 0; i < W.size(); ++i)
    if (can_access[i] > 0 && W[i] == 1.f) { trusted = false; break; }
  if (!trusted)
    for (size_t i = 0; i < W.size(); ++i) W[i] = can_access[i] > 0 ? 2.f : 1.f;
  // everything reachable over NVLink: PCI-E links are not used at all; otherwise they are kept, discounted for
  // the queueing of transfers through the CPU (gpu_topology.h:223-245)
  const bool connected = LinksConnected(W, n);
  for (auto& w : W)
    if (w == 1.f) w = connected ? 0.f : static_cast<float>(1. / n);
  return W;
}

namespace {

// D(v) = (cost of v's links into the other half) - (cost of its links inside its own half), for the two halves
// marked +1 / -1 in `side` (0: not in the cluster being split).  Float accumulation, column order
// (gemv + ewisemult, gpu_topology.h:259-283).
void Gains(const std::vector<float>& W, const std::vector<int>& side, std::vector<float>* D) {
  const int n = static_cast<int>(side.size());
  for (int r = 0; r < n; ++r) {
    float y = 0.f;
    for (int c = 0; c < n; ++c) y += W[r * n + c] * static_cast<float>(side[c]);
    (*D)[r] = y * (-1.f * static_cast<float>(side[r]));
  }
}

}  // namespace

// One Kernighan-Lin bisection of every cluster that has more than two members (gpu_topology.h:326-480).
// `pairs` receives (cluster, new cluster) for a split and (cluster, -size) for a cluster of one or two.
// Returns true when nothing was left to split.
bool BisectClusters(const std::vector<float>& W, std::vector<int>* color, int* n_colors,
                    std::vector<std::pair<int, int>>* pairs, void* engine) {
  std::mt19937& gen = *static_cast<std::mt19937*>(engine);
  const int n = static_cast<int>(color->size());
  const int colors_at_entry = *n_colors;        // clusters created by this pass are not visited by it
  std::vector<int> population(colors_at_entry, 0);
  for (int c : *color) population[c]++;
  std::vector<int> side(n, 0), kept(n, 0);
  std::vector<float> D(n, 0.f);
  bool nothing_to_split = true;

  for (int c = 0; c < colors_at_entry; ++c) {
    const int size = population[c];
    if (size <= 2) { pairs->emplace_back(c, -size); continue; }
    nothing_to_split = false;

    // a random balanced start: shuffle the members, first half one side, the rest the other
    std::vector<int> members;
    for (int v = 0; v < n; ++v) {
      if ((*color)[v] == c) members.push_back(v); else side[v] = 0;
    }
    std::shuffle(members.begin(), members.end(), gen);
    for (int k = 0; k < size; ++k) side[members[k]] = k < size / 2 ? 1 : -1;

    // refinement rounds.  `best` / `best_len` survive from one round to the next (they are only cleared when a
    // round runs out of improving swaps), as in the reference (:388-390, :420-423, :436-444).
    float best = 0.f;
    int best_len = -1;
    unsigned rounds = 0;
    do {
      ++rounds;
      kept = side;
      Gains(W, side, &D);
      std::vector<int> sa, sb;
      std::vector<float> sg;
      std::vector<char> moved(n, 0);
      for (int it = 0; it < size / 2; ++it) {
        // best swap a <-> b over the upper triangle, first of equals; only the row vertex is required to be
        // unmoved (FindBestMove, :293-316)
        int a = -1, b = -1;
        float g = 0.f;
        for (int r = 0; r < n; ++r) {
          if (side[r] == 0 || moved[r]) continue;
          for (int q = r + 1; q < n; ++q) {
            if (side[q] == 0 || side[q] == side[r]) continue;
            const float cost = D[r] + D[q] - 2 * W[r * n + q];
            if (cost > g) { g = cost; a = r; b = q; }
          }
        }
        if (!(g > 0)) { best = 0.f; break; }
        sa.push_back(a); sb.push_back(b); sg.push_back(g);
        side[a] = -side[a];
        side[b] = -side[b];
        moved[a] = moved[b] = 1;
        Gains(W, side, &D);
        D[a] = 0.f;
        D[b] = 0.f;
      }
      for (size_t k = 0; k < sg.size(); ++k) {
        if (k > 0) sg[k] += sg[k - 1];
        if (sg[k] > best) { best = sg[k]; best_len = static_cast<int>(k) + 1; }
      }
      if (best > 0) {
        const int len = std::min<int>(best_len, static_cast<int>(sa.size()));
        for (int i = 0; i < len; ++i) std::swap(kept[sa[i]], kept[sb[i]]);
      }
      side = kept;
    } while (best > 0 && rounds <= static_cast<unsigned>(n));

    for (int v = 0; v < n; ++v)
      if (side[v] == -1) (*color)[v] = *n_colors;
    pairs->emplace_back(c, *n_colors);
    ++*n_colors;
  }
  return nothing_to_split;
}

namespace {

// the state of one root's Kernighan-Lin attempt; Grow() appends one level to the tree
struct Attempt {
  std::vector<int> color;
  int n_colors = 1;
  std::unordered_set<int> roots;       // one vertex per cluster: where the tree enters it
  std::vector<size_t> tree, scan;
  std::vector<std::pair<int, int>> pairs;
};

int RootOf(const std::vector<int>& color, int c, const std::unordered_set<int>& roots) {
  for (int r : roots) if (color[r] == c) return r;
  return -1;
}

// KLGenerateBinaryTree, gpu_topology.h:597-700: for every pair of clusters the bisection has just produced pick
// the heaviest link from the cluster's root into the new cluster (a random one of several equally heavy), then
// append (node, child) for every node of the deepest level.  Returns false when some cluster cannot be reached:
// the attempt is abandoned and restarted with the random engine where it now stands.
bool Grow(const std::vector<float>& W, Attempt* at, std::mt19937* gen) {
  const int n = static_cast<int>(at->color.size());
  std::unordered_set<int> next_roots;
  std::unordered_map<int, int> child_of;
  for (size_t i = 0; i < at->pairs.size(); ++i) {
    if (i == 0) at->scan.push_back(at->tree.size());
    const int first = at->pairs[i].first, second = at->pairs[i].second;
    int parent = -1, child = -1;
    if (second == -2 || second == -1) {
      // a cluster of two: the root takes the other member; of one: the root sits this level out
      parent = RootOf(at->color, first, at->roots);
      if (parent == -1) return false;
      if (second == -2) {
        for (int v = 0; v < n; ++v)
          if (at->color[v] == first && v != parent) { child = v; break; }
      } else {
        child = parent;
      }
    } else {
      int from = first;
      parent = RootOf(at->color, from, at->roots);
      if (parent == -1) { from = second; parent = RootOf(at->color, from, at->roots); }
      if (parent == -1) return false;          // (the reference would index the matrix with -1 here)
      const int dest = from == first ? second : first;
      // FindBestEdge, :567-585: candidates = the columns of the heaviest link, in column order; the initial
      // {-1} only survives while every link seen weighs nothing
      std::vector<int> cand{-1};
      float heaviest = 0.f;
      for (int v = 0; v < n; ++v) {
        if (v == parent || at->color[v] != dest) continue;
        const float w = W[parent * n + v];
        if (w > heaviest) cand.clear();
        if (w >= heaviest) { cand.push_back(v); heaviest = w; }
      }
      if (cand[0] != -1) {
        std::shuffle(cand.begin(), cand.end(), *gen);
        child = cand[0];
      }
      next_roots.insert(parent);
      if (child == -1) return false;
      next_roots.insert(child);
    }
    child_of[parent] = child;
  }
  const size_t levels = at->scan.size();
  const size_t start = at->scan[levels - 2], end = at->scan[levels - 1];
  for (size_t i = start; i < end; ++i) {
    const int node = static_cast<int>(at->tree[i]);
    // the second of two equal neighbours is a node that already sat out the level above
    const int child = (i != start && at->tree[i] == at->tree[i - 1]) ? node : child_of[node];
    at->tree.push_back(node);
    at->tree.push_back(child);
  }
  at->pairs.clear();
  at->roots = std::move(next_roots);
  return true;
}

}  // namespace

// Postprocess, gpu_topology.h:746-770: a GPU that fills several leaves is made to repeat next to itself, level
// by level from the top, so that the reduce sees (g, g) pairs -- nothing to send -- instead of redundant sends.
void FoldRepeats(std::vector<int>* leaves, int n, int depth) {
  std::vector<int>& r = *leaves;
  const int len = static_cast<int>(r.size());
  for (int level = depth - 1; level >= 0; --level) {
    const int stride = 1 << level;
    std::vector<int> above(n, 0), here(n, 0);
    for (int i = 0; i < len; i += 2 * stride) above[r[i]]++;
    for (int i = 0; i < len; i += stride) here[r[i]]++;
    for (int i = len - stride; i - stride >= 0; i -= 2 * stride) {
      const int from = r[i], dest = r[i - stride];
      if ((here[from] > 1 || above[from] >= 1) && from != dest) {
        r[i] = dest;
        here[from]--;
      }
    }
  }
}

// ComputeTreeWeight, gpu_topology.h:778-813: sum of the link weights a tree uses; with `penalty`, -100 for a
// link used twice and -10 for a GPU that receives twice on one level above the leaves.
float TreeWeight(const std::vector<float>& W, const std::vector<int>& r, int n, int depth, bool penalty) {
  float weight = 0.f;
  std::unordered_set<int> used;
  const size_t len = r.size();
  for (int level = 0; level < depth; ++level) {
    const size_t stride = static_cast<size_t>(1) << level;
    std::vector<char> busy(n, 0);
    for (size_t j = 0; j + stride < len; j += 2 * stride) {
      const int from = r[j], dest = r[j + stride];
      if (from != dest) {
        weight += W[from * n + dest];
        if (penalty && used.count(from * n + dest)) weight -= 100;
        used.insert(from * n + dest);
        used.insert(dest * n + from);
      }
      busy[from] = 1;
      if (level > 0 && busy[dest] && penalty) weight -= 10;
      busy[dest] = 1;
    }
  }
  return weight;
}

// IsValid, gpu_topology.h:727-791: can the first `row` leaves of `state` (-1 = not placed) still become a
// balanced binary spanning tree whose every edge is a link?
bool Admissible(const std::vector<float>& W, const std::vector<int>& state, int n, int row, int depth) {
  for (int level = 0; level < depth; ++level) {
    const int stride = 1 << level;
    for (int j = 0; j + stride < row; j += 2 * stride) {
      const int from = state[j], dest = state[j + stride];
      if (W[from * n + dest] == 0.f && from != dest) return false;
    }
  }
  std::vector<char> seen(n, 0);
  int distinct = 0;
  for (int v : state) {
    if (v == -1) continue;
    if (v >= n) return false;
    if (!seen[v]) { seen[v] = 1; ++distinct; }
  }
  const int spare = (1 << depth) - n;       // leaves that must repeat a GPU
  if (row < n) {
    if (distinct > row || distinct < row - spare) return false;
  } else if (row == static_cast<int>(state.size())) {
    if (distinct != n) return false;
  }
  return true;
}

namespace {

// BacktrackGenerateBinaryTree, gpu_topology.h:846-1016: leaves are placed left to right, GPUs tried in
// ascending order; up to 8 GPUs every admissible placement is weighed and the first of the heaviest kept,
// above that the first admissible placement is taken.
bool Search(const std::vector<float>& W, int n, int root, std::vector<size_t>* tree, std::vector<size_t>* scan) {
  tree->clear();
  scan->clear();
  const int depth = TreeDepth(n);
  const int len = 1 << depth;
  const bool exhaustive = depth <= 3;
  std::vector<int> state(len, -1), best(len, -1);
  float best_weight = std::numeric_limits<float>::lowest();
  state[0] = root;
  int row = 1, first = 0;                    // next leaf to place, first GPU still to be tried there
  while (row >= 1 && row < len) {
    int v = first;
    for (; v < n; ++v) {
      state[row] = v;
      if (Admissible(W, state, n, row + 1, depth)) break;
    }
    if (v == n) {                            // nothing fits: undo the previous leaf and move it on
      state[row] = -1;
      --row;
      first = row >= 1 ? state[row] + 1 : 0;
      continue;
    }
    if (row + 1 < len) { ++row; first = 0; continue; }
    std::vector<int> cand = state;
    FoldRepeats(&cand, n, depth);
    const float w = TreeWeight(W, cand, n, depth, true);
    if (w > best_weight) { best_weight = w; best = cand; }
    if (!exhaustive) break;
    first = v + 1;                           // keep going with the next GPU for the last leaf
  }
  for (int v : best) if (v == -1) return false;
  // FormTopology, :833-855: level l of the array form is every 2^(depth-l)-th leaf
  scan->push_back(tree->size());
  for (int l = depth; l > 0; --l) {
    for (int j = 0; j < len; j += 1 << l) tree->push_back(best[j]);
    scan->push_back(tree->size());
  }
  tree->insert(tree->end(), best.begin(), best.end());
  scan->push_back(tree->size());
  return true;
}

// ComputeTreesFromRoot, gpu_topology.h:1018-1098
void TreeFromRoot(std::vector<float>* W, int n, int root, float alpha, bool backtrack, std::vector<size_t>* tree,
                  std::vector<size_t>* scan) {
  std::mt19937 gen(1);
  const Attempt fresh = [&] {
    Attempt a;
    a.color.assign(n, 0);
    a.roots.insert(root);
    a.tree = *tree;
    a.scan = *scan;
    return a;
  }();
  Attempt at;
  bool done = false, restart = true;
  int restarts = 0;
  while (!backtrack && (!done || restart)) {
    if (restart) at = fresh;
    done = BisectClusters(*W, &at.color, &at.n_colors, &at.pairs, &gen);
    restart = !Grow(*W, &at, &gen);
    if (restart && ++restarts > 10) break;
  }
  bool ok = true;
  if (restart) {
    ok = Search(*W, n, root, tree, scan);
  } else {
    *tree = at.tree;
    *scan = at.scan;
    scan->push_back(tree->size());
  }
  MXKV_CHECK(ok) << "MXNET_KVSTORE_USETREE: no balanced binary tree over the GPU links from root " << root;
  // UpdateWeight, :829-841: every link this tree uses is discounted for the trees still to be built
  for (size_t i = 1; i + 1 < tree->size(); i += 2) {
    const size_t parent = (*tree)[i], child = (*tree)[i + 1];
    if (parent != child && parent < static_cast<size_t>(n) * n && child < static_cast<size_t>(n) * n) {
      (*W)[parent * n + child] *= alpha;
      (*W)[child * n + parent] *= alpha;
    }
  }
}

}  // namespace

void ComputeTrees(const std::vector<float>& W, int n, float alpha, bool backtrack, TreeSet* out) {
  MXKV_CHECK(n >= 1 && W.size() == static_cast<size_t>(n) * n) << "link matrix must be " << n << " x " << n;
  std::vector<float> w = W;
  out->n = n;
  out->depth = TreeDepth(n);
  out->topo.assign(n, {});
  out->scan.assign(n, {});
  for (int r = 0; r < n; ++r) {
    out->topo[r].push_back(r);
    out->scan[r].push_back(0);
    TreeFromRoot(&w, n, r, alpha, backtrack, &out->topo[r], &out->scan[r]);
  }
}

// Replays CommDeviceTree::ReduceInner (comm_tree.h:91-177) on expressions instead of arrays.  Each GPU's merge
// buffer starts as its own value; level by level from the leaves a pair (dest, from) with dest != from copies
// from's buffer to dest and dest's buffer becomes (dest + from); the buffer of the root is the result.  The
// expression is then flattened left to right.
ReduceProgram ReduceProgramOf(const std::vector<size_t>& topo, const std::vector<size_t>& scan, int depth, int n) {
  MXKV_CHECK(n >= 1 && n <= kMaxRanks) << "reduce program over " << n << " participants";
  MXKV_CHECK(static_cast<int>(scan.size()) == depth + 2 && scan.back() == topo.size() && !topo.empty())
      << "MXNET_KVSTORE_USETREE: the tree has " << scan.size() << " level marks, expected " << depth + 2
      << " (the reference's reduce would read past its tree here)";
  struct Node { int left, right, leaf; };
  std::vector<Node> nodes;
  std::vector<int> buf(n, -1);                 // expression held by each GPU's merge buffer
  for (size_t j = scan[depth]; j < scan[depth + 1]; ++j) {
    const int g = static_cast<int>(topo[j]);
    MXKV_CHECK(g >= 0 && g < n) << "tree names participant " << g;
    if (buf[g] < 0) { nodes.push_back(Node{-1, -1, g}); buf[g] = static_cast<int>(nodes.size()) - 1; }
  }
  for (int g = 0; g < n; ++g) MXKV_CHECK(buf[g] >= 0) << "participant " << g << " is not a leaf of the tree";
  for (int level = depth; level > 0; --level) {
    std::vector<std::vector<int>> operands(n);
    int dest = 0;
    bool second = false;
    for (size_t j = scan[level]; j < scan[level + 1]; ++j, second = !second) {
      const int g = static_cast<int>(topo[j]);
      if (!second) {
        dest = g;
        if (operands[dest].empty()) operands[dest].push_back(buf[dest]);
      } else if (g != dest) {
        // the one receive buffer of `dest` on this level (copy_buf[..][0], kBranch = 2)
        MXKV_CHECK(operands[dest].size() == 1) << "participant " << dest << " receives twice on level " << level;
        operands[dest].push_back(buf[g]);
      }
    }
    size_t child = scan[level];
    for (size_t i = scan[level - 1]; i < scan[level]; ++i, child += 2) {
      const int g = static_cast<int>(topo[i]);
      if (operands[g].size() > 1 && topo[child] != topo[child + 1]) {
        nodes.push_back(Node{operands[g][0], operands[g][1], -1});
        buf[g] = static_cast<int>(nodes.size()) - 1;
        operands[g].resize(1);
        operands[g][0] = buf[g];
      }
    }
  }
  ReduceProgram rp;
  rp.n = 0;
  int bit = 0, pending = 0;
  std::vector<char> seen(n, 0);
  // iterative post-order: (node, state) with state 0 = visit left, 1 = visit right, 2 = emit the add
  std::vector<std::pair<int, int>> stack{{buf[topo[0]], 0}};
  while (!stack.empty()) {
    auto& top = stack.back();
    const Node nd = nodes[top.first];
    if (nd.leaf >= 0) {
      MXKV_CHECK(!seen[nd.leaf] && rp.n < kMaxRanks) << "participant " << nd.leaf << " enters the sum twice";
      // what is pending now is what the kernel holds in registers while it takes this value
      MXKV_CHECK(pending <= kTreeStack) << "reduction tree deeper than the kernel's " << kTreeStack << " pending sums";
      if (rp.n > 0) rp.prog |= 1u << bit++;            // the previous value's step ends here
      seen[nd.leaf] = 1;
      rp.leaf[rp.n++] = nd.leaf;
      ++pending;
      stack.pop_back();
    } else if (top.second == 0) {
      top.second = 1;
      stack.emplace_back(nd.left, 0);
    } else if (top.second == 1) {
      top.second = 2;
      stack.emplace_back(nd.right, 0);
    } else {
      ++bit;                                           // a 0 bit: one add
      --pending;
      stack.pop_back();
    }
  }
  rp.prog |= 1u << bit++;
  MXKV_CHECK(rp.n == n && bit == 2 * n - 1) << "the tree sums " << rp.n << " of " << n << " participants";
  return rp;
}

std::vector<float> QueryLinkWeights(const std::vector<int>& devs) {
  const int n = static_cast<int>(devs.size());
  if (const char* env = std::getenv("MXKV_B200_TREE_LINKS")) {
    std::vector<float> W;
    std::stringstream ss(env);
    std::string tok;
    while (std::getline(ss, tok, ',')) W.push_back(static_cast<float>(std::atof(tok.c_str())));
    MXKV_CHECK(W.size() == static_cast<size_t>(n) * n) << "MXKV_B200_TREE_LINKS holds " << W.size() << " numbers, this push has "
                                                        << n << " GPUs";
    return W;
  }
  std::vector<int> perf(static_cast<size_t>(n) * n, 0), access(static_cast<size_t>(n) * n, 0);
  for (int r = 0; r < n; ++r) {
    for (int c = 0; c < n; ++c) {
      if (r == c) continue;
      int v = 0;
      if (cudaDeviceGetP2PAttribute(&v, cudaDevP2PAttrPerformanceRank, devs[r], devs[c]) != cudaSuccess) { v = 0; cudaGetLastError(); }
      perf[r * n + c] = v;
      int a = 0;
      if (cudaDeviceCanAccessPeer(&a, devs[r], devs[c]) != cudaSuccess) { a = 0; cudaGetLastError(); }
      access[r * n + c] = a;
    }
  }
  return LinkWeights(n, perf, access);
}

}  // namespace topo
}  // namespace mxkv
