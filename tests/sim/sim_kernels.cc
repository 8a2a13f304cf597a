// sim_kernels.cc -- semantic emulators of the engine's kernels for the simulated CUDA runtime
// (fake_cudart.cc).  Each emulator does, sequentially and on host memory, what the kernel of the same
// name is specified to do with its launch arguments: same descriptors (csrc/kernels.h, norm_kernels.h,
// rsp_kernels.h), same per-element arithmetic (csrc/optim_math.h, csrc/norm_math.h compiled for the
// host), same summation order over the sources.  Cross-GPU rendezvous are not needed: kernels run to
// completion one after the other, which is a valid schedule of the real system because the host never
// lets one rank's stores alias a range another rank still has to read.  Test infrastructure only.
#include <cuda_runtime_api.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <sched.h>
#include <ctime>
#include <unistd.h>
#include <functional>
#include <map>
#include "sim.h"
#include "optim_math.h"
#include "tree_math.h"
#include "norm_math.h"
#include "rsp_kernels.h"

namespace sim {
bool HostEmuTreeLaunch(const void* launch);    // hostemu_tree.cc
bool HostEmuDenseLaunch(const void* launch);   // hostemu_dense.cc
// hostemu_norm.cc
bool HostEmuNormFirst(const void* launch, int grad_only);
bool HostEmuNormMid(const void* launch);
bool HostEmuNormApply(const void* launch);
bool HostEmuNormFinalize(const void* works, const int64_t* prefix, int nworks, int nslots, int s0, int s1, int s2);
void HostEmuSumSq(int type, const void* items, const int64_t* prefix, int nitems, int64_t total_chunks, float scale, float* psum,
                  int chunk_elems, int grid);
void HostEmuSumSqFinalize(const int64_t* prefix, const float* psum, float* out_sumsq, float* out_bad, int nitems);
void HostEmuAllFiniteFlag(const float* bad, int n, float* out, int init);
}

using namespace mxkv;

namespace sim {

// ---- cross-process rendezvous (one simulated GPU per process: MXKV_SIM_MP=1) -------------------------------
// The signal pads live in shared memory there, so the flag protocol of csrc/device_utils.cuh can run for
// real, for "block 0" (an emulated kernel is one sequential block): write the next flag value into every
// rank's pad, spin until every rank has written it into ours.  In the single-process simulation kernels run
// to completion one after another and a rendezvous would only deadlock: it is skipped.
static bool MultiProcess() {
  static const bool mp = getenv("MXKV_SIM_MP") != nullptr;
  return mp;
}

static void Rendezvous(const SyncArgs& s, int offset) {
  if (s.mode == SYNC_NONE || s.world <= 1) return;
  const uint32_t flag = s.epoch;                     // host-provided, the same on every participant
  if (flag == 0) { fprintf(stderr, "[mxkv sim] collective launch without a rendezvous epoch\n"); abort(); }
  if (!MultiProcess()) {
    // single process: kernels run one after another, so nobody can wait -- but the protocol can be CHECKED: the n
    // participants of a collective launch must all carry the same flag value, a value this pad has not seen before
    // (round 1's per-device counters violated the first rule as soon as the set of GPUs changed between launches;
    // the first 8-GPU hardware run dead-locked on it)
    static std::map<const uint32_t*, uint32_t> last_on_pad[2];
    static std::map<uint32_t, std::pair<int, int>> open_launch[2];      // epoch -> (world, participants seen)
    const int which = offset == kSigStartOff ? 0 : 1;
    uint32_t& last = last_on_pad[which][s.self];
    if (flag <= last) { fprintf(stderr, "[mxkv sim] rendezvous epoch %u reused on a pad (last %u)\n", flag, last); abort(); }
    last = flag;
    auto& ol = open_launch[which][flag];
    if (ol.second == 0) ol.first = s.world;
    if (ol.first != s.world) { fprintf(stderr, "[mxkv sim] participants of one launch disagree on its size\n"); abort(); }
    if (++ol.second == s.world) open_launch[which].erase(flag);
    if (open_launch[which].size() > 64) { fprintf(stderr, "[mxkv sim] collective launches left incomplete\n"); abort(); }
    return;
  }
  for (int r = 0; r < s.world; ++r)
    __atomic_store_n(s.peers[r] + offset + s.rank, flag, __ATOMIC_RELEASE);
  for (int r = 0; r < s.world; ++r) {
    const uint32_t* mine = s.self + offset + r;
    // a peer may be busy for a long time (its Python side checks results against the oracle for every
    // rank): wait by the clock, and politely
    long spins = 0;
    const time_t t0 = time(nullptr);
    while (__atomic_load_n(mine, __ATOMIC_ACQUIRE) != flag) {
      if ((++spins & 255) == 0) {
        if (spins > 4096) usleep(100); else sched_yield();
        if (time(nullptr) - t0 > 600) {
          fprintf(stderr, "[mxkv sim] rendezvous timed out (rank %d waits for %d)\n", s.rank, r);
          abort();
        }
      }
    }
  }
}
static void RendezvousStart(const SyncArgs& s) { Rendezvous(s, kSigStartOff); }
static void RendezvousEnd(const SyncArgs& s) { Rendezvous(s, kSigEndOff); }

void ParseName(const std::string& demangled, std::string* base, std::vector<std::string>* targs) {
  std::string s = demangled;
  for (size_t p; (p = s.find("(anonymous namespace)")) != std::string::npos;) s.replace(p, 21, "anon");
  if (s.compare(0, 5, "void ") == 0) s = s.substr(5);
  const size_t paren = s.find('(');
  std::string head = paren == std::string::npos ? s : s.substr(0, paren);
  targs->clear();
  const size_t lt = head.find('<');
  if (lt == std::string::npos) { *base = head; return; }
  *base = head.substr(0, lt);
  std::string inner = head.substr(lt + 1, head.rfind('>') - lt - 1);
  size_t pos = 0;
  int depth = 0;
  std::string cur;
  for (; pos < inner.size(); ++pos) {
    const char c = inner[pos];
    if (c == '<' || c == '(') ++depth;
    if (c == '>' || c == ')') --depth;
    if (c == ',' && depth == 0) { targs->push_back(cur); cur.clear(); continue; }
    if (c == ' ' && cur.empty()) continue;
    cur.push_back(c);
  }
  if (!cur.empty()) targs->push_back(cur);
  for (auto& t : *targs) {          // "(mxkv::OptKind)2" or "2" or "true"
    const size_t rp = t.rfind(')');
    if (rp != std::string::npos && t[0] == '(') t = t.substr(rp + 1);
  }
}

namespace {

// ---- element types ---------------------------------------------------------------------------------
template <typename T> struct H;
template <> struct H<float> {
  static float to(float v) { return v; }
  static float from(float v) { return v; }
};
template <> struct H<__half> {
  static float to(__half v) { return __half2float(v); }
  static __half from(float v) { return __float2half_rn(v); }
};
template <> struct H<__nv_bfloat16> {
  static float to(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __nv_bfloat16 from(float v) { return __float2bfloat16_rn(v); }
};

template <> struct H<double> {
  static float to(double v) { return static_cast<float>(v); }
  static double from(float v) { return static_cast<double>(v); }
};

int ToInt(const std::string& s) { return s == "true" ? 1 : (s == "false" ? 0 : atoi(s.c_str())); }

// gather the n replicas of element e and add them in the reference's association order
template <typename T>
float OrderedSum(const void* const* src, int n, int64_t e, int order, bool native_half_add, uint32_t tree_prog = 0) {
  if (order == ORDER_TREE) {          // the device's own evaluator (csrc/tree_math.h), compiled for the host
    TreeSum<float, 1> sum;
    sum.begin(tree_prog);
    for (int k = 0; k < n; ++k) {
      const float x[1] = {H<T>::to(static_cast<const T*>(src[k])[e])};
      sum.take(x, [native_half_add](float l, float r) {
        float s = __fadd_rn(l, r);
        if (native_half_add) s = H<T>::to(H<T>::from(s));
        return s;
      });
    }
    float y[1];
    sum.result(y);
    return y[0];
  }
  float acc = 0.f, grp = 0.f;
  for (int k = 0; k < n; ++k) {
    const float x = H<T>::to(static_cast<const T*>(src[k])[e]);
    if (k == 0) {
      acc = x;
    } else if (order == ORDER_DEVICE || n <= 2) {
      float s = __fadd_rn(acc, x);
      if (native_half_add) s = H<T>::to(H<T>::from(s));
      acc = s;
    } else {
      const int pos = (k - 1) & 3;
      grp = (pos == 0) ? x : __fadd_rn(grp, x);
      if (pos == 3 || k == n - 1) acc = __fadd_rn(acc, grp);
    }
  }
  return acc;
}

template <typename T, int OPT>
void DenseEntry(const DenseLaunch& L, const TensorWork& tw, bool mp) {
  Hyper h;
  h.lr = tw.lr; h.wd = tw.wd; h.eta = tw.eta;
  h.rescale = L.rescale; h.clip = L.clip; h.momentum = L.momentum;
  h.beta1 = L.beta1; h.beta2 = L.beta2; h.eps = L.eps;
  const bool native_half_add = (sizeof(T) == 2) && !L.fp32_accum && (OPT == OPT_NONE);
  constexpr bool HAS_S0 = OPT == OPT_SGD_MOM || OPT == OPT_ADAM || OPT == OPT_ADAMW || OPT == OPT_ADAM_STD;
  constexpr bool HAS_S1 = OPT == OPT_ADAM || OPT == OPT_ADAMW || OPT == OPT_ADAM_STD;
  for (int64_t e = tw.begin; e < tw.end; ++e) {
    const float acc = OrderedSum<T>(tw.src, tw.n_src, e, L.order, native_half_add, tw.tree_prog);
    float wnew = acc;
    if (OPT != OPT_NONE) {
      const float w = mp ? tw.w32[e] : H<T>::to(static_cast<const T*>(tw.w)[e]);
      float s0 = HAS_S0 ? tw.s0[e] : 0.f, s1 = HAS_S1 ? tw.s1[e] : 0.f;
      wnew = update_one<OPT>(acc, w, s0, s1, h);
      if (HAS_S0) tw.s0[e] = s0;
      if (HAS_S1) tw.s1[e] = s1;
      if (mp) tw.w32[e] = wnew;
    }
    for (int j = 0; j < tw.n_out; ++j) static_cast<T*>(tw.out[j])[e] = H<T>::from(wnew);
  }
}

template <typename T>
bool DenseT(const DenseLaunch& L, int opt, bool mp) {
  RendezvousStart(L.sync);
  struct AtExit { const SyncArgs& s; ~AtExit() { RendezvousEnd(s); } } at_exit{L.sync};
  for (int i = 0; i < L.nworks; ++i) {
    const TensorWork& tw = L.works[i];
    switch (opt) {
      case OPT_NONE: DenseEntry<T, OPT_NONE>(L, tw, mp); break;
      case OPT_SGD: DenseEntry<T, OPT_SGD>(L, tw, mp); break;
      case OPT_SGD_MOM: DenseEntry<T, OPT_SGD_MOM>(L, tw, mp); break;
      case OPT_ADAM: DenseEntry<T, OPT_ADAM>(L, tw, mp); break;
      case OPT_ADAMW: DenseEntry<T, OPT_ADAMW>(L, tw, mp); break;
      case OPT_TEST: DenseEntry<T, OPT_TEST>(L, tw, mp); break;
      case OPT_SGD_STD: DenseEntry<T, OPT_SGD_STD>(L, tw, mp); break;
      case OPT_ADAM_STD: DenseEntry<T, OPT_ADAM_STD>(L, tw, mp); break;
      default: return false;
    }
  }
  return true;
}

// The per-thread, staged and typed-sum kernels run from their OWN source (hostemu_dense.cc: csrc/kernels.cu compiled
// for the host) unless MXKV_SIM_DENSE=semantic asks for the independent emulators.  The cross-GPU rendezvous stays here.
bool DenseFromSource(const DenseLaunch& L) {
  static const bool semantic = [] { const char* v = getenv("MXKV_SIM_DENSE"); return v != nullptr && std::string(v) == "semantic"; }();
  if (semantic) return false;
  if (MultiProcess() && L.sync.mode != SYNC_NONE) {        // (as in TreeFromSource: the kernels' own rendezvous)
    if (!::sim::HostEmuDenseLaunch(&L)) { fprintf(stderr, "sim: kernels.cu's launcher refused the launch\n"); abort(); }
    return true;
  }
  RendezvousStart(L.sync);
  struct AtExit { const SyncArgs& s; ~AtExit() { RendezvousEnd(s); } } at_exit{L.sync};
  DenseLaunch local = L;
  local.sync.mode = SYNC_NONE;
  if (!::sim::HostEmuDenseLaunch(&local)) { fprintf(stderr, "sim: kernels.cu's launcher refused the launch\n"); abort(); }
  return true;
}

bool Dense(const std::vector<std::string>& t, void** args) {          // kv_dense_kernel<T, OPT, MP, SMALLN>
  const DenseLaunch& L = *static_cast<const DenseLaunch*>(args[0]);
  if (L.order != ORDER_TREE && DenseFromSource(L)) return true;
  const int opt = ToInt(t[1]);
  const bool mp = ToInt(t[2]) != 0;
  if (t[0] == "float") return DenseT<float>(L, opt, mp);
  if (t[0] == "__half") return DenseT<__half>(L, opt, mp);
  if (t[0] == "__nv_bfloat16") return DenseT<__nv_bfloat16>(L, opt, mp);
  return false;
}

// The tree-order kernels run from their OWN source (hostemu_tree.cc: csrc/tree_kernels.cu compiled for the host)
// unless MXKV_SIM_TREE=semantic asks for the independent emulators below.  The cross-GPU rendezvous stays here.
bool TreeFromSource(const DenseLaunch& L) {
  static const bool semantic = [] { const char* v = getenv("MXKV_SIM_TREE"); return v != nullptr && std::string(v) == "semantic"; }();
  if (semantic) return false;
  if (MultiProcess() && L.sync.mode != SYNC_NONE) {
    // one process per simulated GPU: the pads are shared memory and every rank runs the same grid block by block, so
    // the kernel's OWN rendezvous (device_utils.cuh: barrier_start / barrier_end, every block's flag slots) runs for real
    if (!::sim::HostEmuTreeLaunch(&L)) { fprintf(stderr, "sim: the tree kernel's launcher refused the launch\n"); abort(); }
    return true;
  }
  RendezvousStart(L.sync);
  struct AtExit { const SyncArgs& s; ~AtExit() { RendezvousEnd(s); } } at_exit{L.sync};
  DenseLaunch local = L;
  local.sync.mode = SYNC_NONE;
  if (!::sim::HostEmuTreeLaunch(&local)) { fprintf(stderr, "sim: the tree kernel's launcher refused the launch\n"); abort(); }
  return true;
}

bool DenseTree(const std::vector<std::string>& t, void** args) {      // kv_dense_tree_kernel<T, OPT, MP>
  const DenseLaunch& L = *static_cast<const DenseLaunch*>(args[0]);
  if (L.order != ORDER_TREE || L.bulk || L.nvls) return false;
  if (TreeFromSource(L)) return true;
  return Dense(t, args);
}

bool SumTreeF64(void** args) {                                         // kv_sum_tree_f64_kernel
  const DenseLaunch& L = *static_cast<const DenseLaunch*>(args[0]);
  if (L.order != ORDER_TREE || L.dtype != kFloat64) return false;
  if (TreeFromSource(L)) return true;
  RendezvousStart(L.sync);
  struct AtExit { const SyncArgs& s; ~AtExit() { RendezvousEnd(s); } } at_exit{L.sync};
  for (int i = 0; i < L.nworks; ++i) {
    const TensorWork& tw = L.works[i];
    for (int64_t e = tw.begin; e < tw.end; ++e) {
      TreeSum<double, 1> sum;
      sum.begin(tw.tree_prog);
      for (int k = 0; k < tw.n_src; ++k) {
        const double x[1] = {static_cast<const double*>(tw.src[k])[e]};
        sum.take(x, [](double l, double r) { return l + r; });
      }
      double y[1];
      sum.result(y);
      for (int j = 0; j < tw.n_out; ++j) static_cast<double*>(tw.out[j])[e] = y[0];
    }
  }
  return true;
}

bool DenseBulk(const std::vector<std::string>& t, void** args) {      // kv_dense_bulk_kernel<OPT, MP>
  const DenseLaunch& L = *static_cast<const DenseLaunch*>(args[0]);
  if (DenseFromSource(L)) return true;
  return DenseT<float>(L, ToInt(t[0]), ToInt(t[1]) != 0);
}

template <typename T>
void TypedSum(const DenseLaunch& L) {
  RendezvousStart(L.sync);
  struct AtExit { const SyncArgs& s; ~AtExit() { RendezvousEnd(s); } } at_exit{L.sync};
  for (int i = 0; i < L.nworks; ++i) {
    const TensorWork& tw = L.works[i];
    for (int64_t e = tw.begin; e < tw.end; ++e) {
      T acc = static_cast<const T*>(tw.src[0])[e];
      for (int k = 1; k < tw.n_src; ++k) acc = static_cast<T>(acc + static_cast<const T*>(tw.src[k])[e]);
      for (int j = 0; j < tw.n_out; ++j) static_cast<T*>(tw.out[j])[e] = acc;
    }
  }
}

bool SumTyped(const std::vector<std::string>& t, void** args) {       // kv_sum_typed_kernel<T>
  const DenseLaunch& L = *static_cast<const DenseLaunch*>(args[0]);
  if (DenseFromSource(L)) return true;
  const std::string& n = t[0];
  if (n == "double") TypedSum<double>(L);
  else if (n == "int") TypedSum<int32_t>(L);
  else if (n == "long") TypedSum<int64_t>(L);
  else if (n == "unsigned char") TypedSum<uint8_t>(L);
  else if (n == "signed char" || n == "char") TypedSum<int8_t>(L);
  else return false;
  return true;
}

bool CastF32(const std::vector<std::string>& t, void** args) {        // kv_cast_f32_kernel<T>(src, dst, n)
  const void* src = *static_cast<const void* const*>(args[0]);
  float* dst = *static_cast<float* const*>(args[1]);
  const int64_t n = *static_cast<const int64_t*>(args[2]);
  for (int64_t i = 0; i < n; ++i) {
    if (t[0] == "float") dst[i] = static_cast<const float*>(src)[i];
    else if (t[0] == "__half") dst[i] = __half2float(static_cast<const __half*>(src)[i]);
    else if (t[0] == "__nv_bfloat16") dst[i] = __bfloat162float(static_cast<const __nv_bfloat16*>(src)[i]);
    else return false;
  }
  return true;
}

bool Quantize(const std::vector<std::string>& t, void** args) {       // kv_quantize_kernel<BITS>
  const float* grad = *static_cast<const float* const*>(args[0]);
  float* residual = *static_cast<float* const*>(args[1]);
  uint32_t* out = *static_cast<uint32_t* const*>(args[2]);
  const int64_t n = *static_cast<const int64_t*>(args[3]);
  const float thr = *static_cast<const float*>(args[4]);
  const int bits = ToInt(t[0]);
  const int64_t nwords = (n + 32 / bits - 1) / (32 / bits);
  for (int64_t w = 0; w < nwords; ++w)
    out[w] = bits == 2 ? quantize_word<2>(grad, residual, n, thr, w) : quantize_word<1>(grad, residual, n, thr, w);
  return true;
}

bool Dequantize(const std::vector<std::string>& t, void** args) {     // kv_dequantize_kernel<BITS>
  const uint32_t* in = *static_cast<const uint32_t* const*>(args[0]);
  float* out = *static_cast<float* const*>(args[1]);
  const int64_t n = *static_cast<const int64_t*>(args[2]);
  const float thr = *static_cast<const float*>(args[3]);
  const int bits = ToInt(t[0]);
  for (int64_t i = 0; i < n; ++i) out[i] = bits == 2 ? dequantize_value<2>(in, thr, i) : dequantize_value<1>(in, thr, i);
  return true;
}

// The reduction shape of block_sums (csrc/norm_kernels.cu): per-thread partials (filled by the caller in the
// kernel's thread-strided order), xor-shuffle tree inside each warp, then the warps in order.  Keeps the
// emulated sums as accurate as the real ones (a sequential float sum over a chunk would not be).
constexpr int kSimThreads = 512;
float BlockTree(float* part, int threads = kSimThreads) {
  for (int w = 0; w < threads / 32; ++w) {
    float* v = part + w * 32;
    for (int off = 16; off >= 1; off >>= 1) {
      float t[32];
      for (int l = 0; l < 32; ++l) t[l] = v[l] + v[l ^ off];
      for (int l = 0; l < 32; ++l) v[l] = t[l];
    }
  }
  float s = 0.f;
  for (int w = 0; w < threads / 32; ++w) s += part[w * 32];
  return s;
}
// thread that handles element e of a chunk starting at cb when the vector path (4 elements per thread
// and step) is taken up to `vec_end` and the scalar tail after it
inline int OwnerThread(int64_t e, int64_t cb, int64_t vec_end) {
  return e < vec_end ? static_cast<int>(((e - cb) / 4) % kSimThreads) : static_cast<int>((e - vec_end) % kSimThreads);
}

// ---- layer-wise adaptive optimizers (csrc/norm_kernels.cu) ---------------------------------------------
float BadTotal(const NormLaunch& L) {
  float t = 0.f;
  for (int i = 0; i < L.n_bad; ++i) t += *L.bad_list[i];
  return t;
}

template <typename T>
bool NormFirstT(const NormLaunch& L, bool mp, bool grad_only) {
  RendezvousStart(L.sync);
  struct AtExit { const SyncArgs& s; ~AtExit() { RendezvousEnd(s); } } at_exit{L.sync};
  for (int i = 0; i < L.nworks; ++i) {
    const NormWork& tw = L.works[i];
    const int64_t nchunks = L.chunk_prefix[i + 1] - L.chunk_prefix[i];
    for (int64_t ci = 0; ci < nchunks; ++ci) {
      const int64_t cb = tw.begin + ci * L.chunk_elems;
      const int64_t ce = std::min<int64_t>(cb + L.chunk_elems, tw.end);
      static thread_local float part[3][kSimThreads];
      std::memset(part, 0, sizeof(part));
      const int64_t vec_end = (tw.flags & 1) ? cb + (ce - cb) / 4 * 4 : cb;
      for (int64_t e = cb; e < ce; ++e) {
        const int th = OwnerThread(e, cb, vec_end);
        const float g = OrderedSum<T>(tw.src, tw.n_src, e, L.order, false);
        const float w = (mp && !grad_only) ? tw.w32[e] : H<T>::to(static_cast<const T*>(tw.w)[e]);
        part[2][th] += not_finite(g) ? 1.0f : 0.0f;
        if (grad_only) {
          tw.aux0[e] = g;
          part[0][th] += w * w;
          const float gs = (L.rescale != 1.0f) ? __fmul_rn(g, L.rescale) : g;
          part[1][th] += gs * gs;
        } else {
          float m = tw.s0[e], v = tw.s1[e];
          const float gh = lamb_step1(g, w, m, v, L, tw);
          tw.s0[e] = m; tw.s1[e] = v; tw.aux0[e] = gh;
          part[0][th] += w * w;
          part[1][th] += gh * gh;
        }
      }
      for (int j = 0; j < 3; ++j) tw.psum[ci * kPsumStride + j] = BlockTree(part[j]);
    }
  }
  return true;
}

// The layer-wise-optimizer kernels run from their OWN source with their real thread counts (hostemu_norm.cc) unless
// MXKV_SIM_NORM=semantic asks for the independent emulators.  phase: 0 first, 1 mid, 2 apply.
bool NormFromSource() {
  static const bool semantic = [] { const char* v = getenv("MXKV_SIM_NORM"); return v != nullptr && std::string(v) == "semantic"; }();
  return !semantic;
}
bool NormLaunchFromSource(const NormLaunch& L, int phase, int grad_only) {
  if (!NormFromSource()) return false;
  NormLaunch local = L;
  const bool own_rendezvous = MultiProcess() && L.sync.mode != SYNC_NONE;   // (as for the dense kernels)
  if (!own_rendezvous) { RendezvousStart(L.sync); local.sync.mode = SYNC_NONE; }
  const bool ok = phase == 0 ? ::sim::HostEmuNormFirst(&local, grad_only)
                             : (phase == 1 ? ::sim::HostEmuNormMid(&local) : ::sim::HostEmuNormApply(&local));
  if (!own_rendezvous) RendezvousEnd(L.sync);
  if (!ok) { fprintf(stderr, "sim: norm_kernels.cu's launcher refused the launch\n"); abort(); }
  return true;
}

bool NormFirst(const std::vector<std::string>& t, void** args) {      // kv_norm_first_kernel<T, MP, GRAD_ONLY>
  const NormLaunch& L = *static_cast<const NormLaunch*>(args[0]);
  const bool mp = ToInt(t[1]) != 0, go = ToInt(t[2]) != 0;
  if (NormLaunchFromSource(L, 0, go ? 1 : 0)) return true;
  if (t[0] == "float") return NormFirstT<float>(L, mp, go);
  if (t[0] == "__half") return NormFirstT<__half>(L, mp, go);
  if (t[0] == "__nv_bfloat16") return NormFirstT<__nv_bfloat16>(L, mp, go);
  return false;
}

bool NormFinalize(const LaunchInfo& info, void** args) {   // (works, prefix, nslots, slot0, slot1, slot2), grid = entries
  const NormWork* works = *static_cast<const NormWork* const*>(args[0]);
  const int64_t* prefix = *static_cast<const int64_t* const*>(args[1]);
  const int nslots = *static_cast<const int*>(args[2]);
  const int slots[3] = {*static_cast<const int*>(args[3]), *static_cast<const int*>(args[4]),
                        *static_cast<const int*>(args[5])};
  if (NormFromSource()) {
    if (!::sim::HostEmuNormFinalize(works, prefix, static_cast<int>(info.grid), nslots, slots[0], slots[1], slots[2])) abort();
    return true;
  }
  for (unsigned b = 0; b < info.grid; ++b) {
    const NormWork& w = works[b];
    const int64_t nchunks = prefix[b + 1] - prefix[b];
    for (int j = 0; j < nslots; ++j) {
      float part[256] = {0.f};
      for (int64_t i = 0; i < nchunks; ++i) part[i % 256] += w.psum[i * kPsumStride + j];
      w.nrm[slots[j]] = BlockTree(part, 256);
    }
  }
  return true;
}

template <typename T>
bool NormMidT(const NormLaunch& L, bool mp, int kind) {
  RendezvousStart(L.sync);
  struct AtExit { const SyncArgs& s; ~AtExit() { RendezvousEnd(s); } } at_exit{L.sync};
  if (L.skip_nonfinite && BadTotal(L) > 0.f) return true;
  for (int i = 0; i < L.nworks; ++i) {
    const NormWork& tw = L.works[i];
    const float g_norm = kind == NORM_LANS ? __fsqrt_rn(rank_total(tw, kNrmG)) : 1.0f;
    const int64_t nchunks = L.chunk_prefix[i + 1] - L.chunk_prefix[i];
    for (int64_t ci = 0; ci < nchunks; ++ci) {
      const int64_t cb = tw.begin + ci * L.chunk_elems;
      const int64_t ce = std::min<int64_t>(cb + L.chunk_elems, tw.end);
      static thread_local float part[3][kSimThreads];
      std::memset(part, 0, sizeof(part));
      const int64_t vec_end = (tw.flags & 1) ? cb + (ce - cb) / 4 * 4 : cb;
      for (int64_t e = cb; e < ce; ++e) {
        const int th = OwnerThread(e, cb, vec_end);
        const float g = tw.aux0[e];
        const float w = mp ? tw.w32[e] : H<T>::to(static_cast<const T*>(tw.w)[e]);
        float m = tw.s0[e], v = tw.s1[e];
        if (kind == NORM_LAMB) {
          const float gh = lamb_step1(g, w, m, v, L, tw);
          tw.aux0[e] = gh;
          part[0][th] += w * w;
          part[1][th] += gh * gh;
        } else {
          float tm, tg;
          lans_step1(g, w, m, v, g_norm, L, tw, tm, tg);
          tw.aux1[e] = tm; tw.aux0[e] = tg;
          part[0][th] += tm * tm;
          part[1][th] += tg * tg;
        }
        tw.s0[e] = m; tw.s1[e] = v;
      }
      for (int j = 0; j < 3; ++j) tw.psum[ci * kPsumStride + j] = BlockTree(part[j]);
    }
  }
  return true;
}

bool NormMid(const std::vector<std::string>& t, void** args) {        // kv_norm_mid_kernel<T, MP, KIND>
  const NormLaunch& L = *static_cast<const NormLaunch*>(args[0]);
  const bool mp = ToInt(t[1]) != 0;
  const int kind = ToInt(t[2]);
  if (NormLaunchFromSource(L, 1, 0)) return true;
  if (t[0] == "float") return NormMidT<float>(L, mp, kind);
  if (t[0] == "__half") return NormMidT<__half>(L, mp, kind);
  if (t[0] == "__nv_bfloat16") return NormMidT<__nv_bfloat16>(L, mp, kind);
  return false;
}

template <typename T>
bool NormApplyT(const NormLaunch& L, bool mp, int flavor) {
  RendezvousStart(L.sync);
  struct AtExit { const SyncArgs& s; ~AtExit() { RendezvousEnd(s); } } at_exit{L.sync};
  bool skip = false;
  if (L.skip_nonfinite) {
    skip = BadTotal(L) > 0.f;
    if (skip && L.overflow_flag != nullptr) *L.overflow_flag = 1;
  }
  for (int i = 0; i < L.nworks; ++i) {
    const NormWork& tw = L.works[i];
    float sc[2] = {0.f, 0.f};
    if (!skip) {
      if (flavor == APPLY_LAMB) apply_scalars<APPLY_LAMB>(tw, L, sc);
      else if (flavor == APPLY_LANS) apply_scalars<APPLY_LANS>(tw, L, sc);
      else apply_scalars<APPLY_LARS>(tw, L, sc);
    }
    for (int64_t e = tw.begin; e < tw.end; ++e) {
      const float w = mp ? tw.w32[e] : H<T>::to(static_cast<const T*>(tw.w)[e]);
      float wn = w;
      if (skip) {
      } else if (flavor == APPLY_LAMB) {
        wn = __fsub_rn(w, __fmul_rn(sc[0], tw.aux0[e]));
      } else if (flavor == APPLY_LANS) {
        wn = __fsub_rn(w, __fadd_rn(__fmul_rn(sc[0], tw.aux1[e]), __fmul_rn(sc[1], tw.aux0[e])));
      } else {
        Hyper h;
        h.lr = sc[0]; h.wd = tw.wd; h.eta = 1.f;
        h.rescale = L.rescale; h.clip = L.clip; h.momentum = L.momentum;
        h.beta1 = 0.f; h.beta2 = 0.f; h.eps = 0.f;
        float unused = 0.f;
        if (flavor == APPLY_LARS_MOM) {
          float mom = tw.s0[e];
          wn = update_one<OPT_SGD_MOM>(tw.aux0[e], w, mom, unused, h);
          tw.s0[e] = mom;
        } else {
          wn = update_one<OPT_SGD>(tw.aux0[e], w, unused, unused, h);
        }
      }
      if (mp && !skip) tw.w32[e] = wn;
      for (int j = 0; j < tw.n_out; ++j) static_cast<T*>(tw.out[j])[e] = H<T>::from(wn);
    }
  }
  return true;
}

bool NormApply(const std::vector<std::string>& t, void** args) {      // kv_norm_apply_kernel<T, MP, FLAVOR>
  const NormLaunch& L = *static_cast<const NormLaunch*>(args[0]);
  const bool mp = ToInt(t[1]) != 0;
  const int flavor = ToInt(t[2]);
  if (NormLaunchFromSource(L, 2, 0)) return true;
  if (t[0] == "float") return NormApplyT<float>(L, mp, flavor);
  if (t[0] == "__half") return NormApplyT<__half>(L, mp, flavor);
  if (t[0] == "__nv_bfloat16") return NormApplyT<__nv_bfloat16>(L, mp, flavor);
  return false;
}

template <typename T>
void SumSqT(const SumSqItem* items, const int64_t* prefix, int nitems, float scale, float* psum, int chunk) {
  for (int it = 0; it < nitems; ++it) {
    const T* x = static_cast<const T*>(items[it].ptr);
    for (int64_t c = prefix[it]; c < prefix[it + 1]; ++c) {
      const int64_t cb = (c - prefix[it]) * chunk;
      const int64_t ce = std::min<int64_t>(cb + chunk, items[it].n);
      static thread_local float part[2][kSimThreads];
      std::memset(part, 0, sizeof(part));
      for (int64_t i = cb; i < ce; ++i) {
        const int th = static_cast<int>((i - cb) % kSimThreads);
        float v = H<T>::to(x[i]);
        part[1][th] += not_finite(v) ? 1.0f : 0.0f;
        if (scale != 1.0f) v = __fmul_rn(v, scale);
        part[0][th] += v * v;
      }
      psum[c * 2] = BlockTree(part[0]); psum[c * 2 + 1] = BlockTree(part[1]);
    }
  }
}

bool SumSq(const std::vector<std::string>& t, void** args, unsigned sumsq_grid) {   // (items, prefix, nitems, total_chunks, scale, psum, chunk)
  const SumSqItem* items = *static_cast<const SumSqItem* const*>(args[0]);
  const int64_t* prefix = *static_cast<const int64_t* const*>(args[1]);
  const int nitems = *static_cast<const int*>(args[2]);
  const float scale = *static_cast<const float*>(args[4]);
  float* psum = *static_cast<float* const*>(args[5]);
  const int chunk = *static_cast<const int*>(args[6]);
  if (NormFromSource()) {
    const int64_t total_chunks = *static_cast<const int64_t*>(args[3]);
    const int type = t[0] == "float" ? 0 : (t[0] == "__half" ? 1 : (t[0] == "__nv_bfloat16" ? 2 : (t[0] == "double" ? 3 : -1)));
    if (type < 0) return false;
    ::sim::HostEmuSumSq(type, items, prefix, nitems, total_chunks, scale, psum, chunk, static_cast<int>(sumsq_grid));
    return true;
  }
  if (t[0] == "float") SumSqT<float>(items, prefix, nitems, scale, psum, chunk);
  else if (t[0] == "__half") SumSqT<__half>(items, prefix, nitems, scale, psum, chunk);
  else if (t[0] == "__nv_bfloat16") SumSqT<__nv_bfloat16>(items, prefix, nitems, scale, psum, chunk);
  else if (t[0] == "double") SumSqT<double>(items, prefix, nitems, scale, psum, chunk);
  else return false;
  return true;
}

bool SumSqFinalize(const LaunchInfo& info, void** args) {   // (prefix, psum, out_sumsq, out_bad), grid = items
  const int64_t* prefix = *static_cast<const int64_t* const*>(args[0]);
  const float* psum = *static_cast<const float* const*>(args[1]);
  float* out_sumsq = *static_cast<float* const*>(args[2]);
  float* out_bad = *static_cast<float* const*>(args[3]);
  if (NormFromSource()) { ::sim::HostEmuSumSqFinalize(prefix, psum, out_sumsq, out_bad, static_cast<int>(info.grid)); return true; }
  for (unsigned b = 0; b < info.grid; ++b) {
    float ps[256] = {0.f}, pb[256] = {0.f};
    for (int64_t i = prefix[b]; i < prefix[b + 1]; ++i) {
      ps[(i - prefix[b]) % 256] += psum[i * 2];
      pb[(i - prefix[b]) % 256] += psum[i * 2 + 1];
    }
    if (out_sumsq) out_sumsq[b] = BlockTree(ps, 256);
    if (out_bad) out_bad[b] = BlockTree(pb, 256);
  }
  return true;
}

bool AllFiniteFlag(void** args) {   // (bad, n, out, init)
  const float* bad = *static_cast<const float* const*>(args[0]);
  const int n = *static_cast<const int*>(args[1]);
  float* out = *static_cast<float* const*>(args[2]);
  const int init = *static_cast<const int*>(args[3]);
  if (NormFromSource()) { ::sim::HostEmuAllFiniteFlag(bad, n, out, init); return true; }
  float v = init ? 1.0f : out[0];
  for (int i = 0; i < n; ++i) if (bad[i] > 0.f) v = 0.0f;
  out[0] = v;
  return true;
}

}  // namespace

bool DispatchRsp(const LaunchInfo& info, const std::string& base, const std::vector<std::string>& t, void** args);

bool Dispatch(const LaunchInfo& info, void** args) {
  std::string base;
  std::vector<std::string> t;
  ParseName(info.name, &base, &t);
  if (base == "mxkv::kv_dense_kernel") return Dense(t, args);
  if (base == "mxkv::kv_dense_bulk_kernel") return DenseBulk(t, args);
  if (base == "mxkv::kv_dense_nvls_kernel")        // (from source only: multimem resolved by fake_driver.cc)
    return DenseFromSource(*static_cast<const DenseLaunch*>(args[0]));
  if (base == "mxkv::kv_dense_tree_kernel") return DenseTree(t, args);
  if (base == "mxkv::kv_sum_tree_f64_kernel") return SumTreeF64(args);
  if (base == "mxkv::kv_sum_typed_kernel") return SumTyped(t, args);
  if (base == "mxkv::kv_cast_f32_kernel") return CastF32(t, args);
  if (base == "mxkv::kv_quantize_kernel") return Quantize(t, args);
  if (base == "mxkv::kv_dequantize_kernel") return Dequantize(t, args);
  if (base == "mxkv::kv_barrier_kernel") {                     // block-0 rendezvous between non-collective kernels
    const SyncArgs& sy = *static_cast<const SyncArgs*>(args[0]);
    RendezvousStart(sy);
    RendezvousEnd(sy);
    return true;
  }
  if (base == "mxkv::kv_norm_first_kernel") return NormFirst(t, args);
  if (base == "mxkv::kv_norm_finalize_kernel") return NormFinalize(info, args);
  if (base == "mxkv::kv_norm_mid_kernel") return NormMid(t, args);
  if (base == "mxkv::kv_norm_apply_kernel") return NormApply(t, args);
  if (base == "mxkv::kv_sumsq_kernel") return SumSq(t, args, info.grid);
  if (base == "mxkv::kv_sumsq_finalize_kernel") return SumSqFinalize(info, args);
  if (base == "mxkv::kv_all_finite_flag_kernel") return AllFiniteFlag(args);
  return DispatchRsp(info, base, t, args);
}

}  // namespace sim
