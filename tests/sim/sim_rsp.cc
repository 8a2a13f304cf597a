// sim_rsp.cc -- emulators of the row_sparse kernels (csrc/rsp_kernels.cu) for the simulated runtime.
#include <cuda_runtime_api.h>
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <ctime>
#include <sched.h>
#include <unistd.h>
#include "sim.h"
#include "rsp_math.h"

using namespace mxkv;

namespace sim {
namespace {

int64_t LowerBound(const int64_t* a, int64_t n, int64_t x) { return std::lower_bound(a, a + n, x) - a; }

template <typename A> const A& Arg(void** args, int i) { return *static_cast<const A*>(args[i]); }

bool First(void** args) {      // (RspSources S, int32_t* first, int64_t cap)
  const RspSources& S = Arg<RspSources>(args, 0);
  int32_t* first = Arg<int32_t*>(args, 1);
  const int64_t cap = Arg<int64_t>(args, 2);
  for (int s = 0; s < S.n; ++s) {
    const int64_t nnz = *S.nnz[s];
    for (int64_t r = 0; r < nnz; ++r) {
      const int64_t id = S.idx[s][r];
      int f = 1;
      for (int t = 0; t < s && f; ++t) {
        const int64_t nt = *S.nnz[t];
        const int64_t p = LowerBound(S.idx[t], nt, id);
        if (p < nt && S.idx[t][p] == id) f = 0;
      }
      first[s * cap + r] = f;
    }
  }
  return true;
}

bool Scan(void** args) {       // (S, first, pf, cap)
  const RspSources& S = Arg<RspSources>(args, 0);
  const int32_t* first = Arg<const int32_t*>(args, 1);
  int32_t* pf = Arg<int32_t*>(args, 2);
  const int64_t cap = Arg<int64_t>(args, 3);
  for (int s = 0; s < S.n; ++s) {
    const int64_t nnz = *S.nnz[s];
    int32_t run = 0;
    for (int64_t i = 0; i < nnz; ++i) { pf[s * (cap + 1) + i] = run; run += first[s * cap + i]; }
    pf[s * (cap + 1) + nnz] = run;
  }
  return true;
}

bool Rank(void** args) {       // (S, first, pf, cap, out_idx, d_nnz_out)
  const RspSources& S = Arg<RspSources>(args, 0);
  const int32_t* first = Arg<const int32_t*>(args, 1);
  const int32_t* pf = Arg<const int32_t*>(args, 2);
  const int64_t cap = Arg<int64_t>(args, 3);
  int64_t* out_idx = Arg<int64_t*>(args, 4);
  int64_t* d_nnz_out = Arg<int64_t*>(args, 5);
  int64_t tot = 0;
  for (int t = 0; t < S.n; ++t) tot += pf[t * (cap + 1) + *S.nnz[t]];
  *d_nnz_out = tot;
  for (int s = 0; s < S.n; ++s) {
    const int64_t nnz = *S.nnz[s];
    for (int64_t r = 0; r < nnz; ++r) {
      if (!first[s * cap + r]) continue;
      const int64_t id = S.idx[s][r];
      int64_t rank = 0;
      for (int t = 0; t < S.n; ++t) {
        const int64_t p = (t == s) ? r : LowerBound(S.idx[t], *S.nnz[t], id);
        rank += pf[t * (cap + 1) + p];
      }
      out_idx[rank] = id;
    }
  }
  return true;
}

template <int OPT>
void RowsT(const RspSources& S, const RspRowArgs& A) {
  const int64_t nrows = *A.d_nnz_out;
  const int64_t L = A.row_len;
  for (int64_t j = 0; j < nrows; ++j) {
    const int64_t id = A.out_idx[j];
    int64_t pos[kMaxSrc];
    for (int t = 0; t < S.n; ++t) {
      const int64_t nt = *S.nnz[t];
      const int64_t p = LowerBound(S.idx[t], nt, id);
      pos[t] = (p < nt && S.idx[t][p] == id) ? p : -1;
    }
    for (int64_t c = 0; c < L; ++c) {
      float acc = 0.f;                                  // set_zero(out) then += inputs in order
      for (int t = 0; t < S.n; ++t)
        if (pos[t] >= 0) acc = __fadd_rn(acc, S.val[t][pos[t] * L + c]);
      if (A.out_val != nullptr) A.out_val[j * L + c] = acc;
      if (OPT != OPT_NONE || A.assign) {
        float* w = A.table + id * L + c;
        if (OPT == OPT_NONE) *w = acc;
        else *w = rsp_lazy_update<(OPT == OPT_NONE ? OPT_SGD : OPT)>(acc, *w, id * L + c, A);
      }
    }
  }
}

bool Rows(const std::vector<std::string>& t, void** args) {   // rsp_rows_kernel<OPT, VEC>(S, A)
  const RspSources& S = Arg<RspSources>(args, 0);
  const RspRowArgs& A = Arg<RspRowArgs>(args, 1);
  switch (atoi(t[0].c_str())) {
    case OPT_NONE: RowsT<OPT_NONE>(S, A); break;
    case OPT_SGD: RowsT<OPT_SGD>(S, A); break;
    case OPT_SGD_MOM: RowsT<OPT_SGD_MOM>(S, A); break;
    case OPT_ADAM: RowsT<OPT_ADAM>(S, A); break;
    default: return false;
  }
  return true;
}

void SortedUnique(std::vector<int64_t> v, int64_t* out, int64_t* d_count) {
  std::sort(v.begin(), v.end());
  v.erase(std::unique(v.begin(), v.end()), v.end());
  if (!v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(int64_t));
  *d_count = static_cast<int64_t>(v.size());
}

bool Unique(void** args) {     // (in, n, out, d_count)
  const int64_t* in = Arg<const int64_t*>(args, 0);
  const int64_t n = Arg<int64_t>(args, 1);
  SortedUnique(std::vector<int64_t>(in, in + n), Arg<int64_t*>(args, 2), Arg<int64_t*>(args, 3));
  return true;
}

bool Pad(void** args) {        // (in, n, buf, npad)
  const int64_t* in = Arg<const int64_t*>(args, 0);
  const int64_t n = Arg<int64_t>(args, 1);
  int64_t* buf = Arg<int64_t*>(args, 2);
  const int64_t npad = Arg<int64_t>(args, 3);
  for (int64_t i = 0; i < npad; ++i) buf[i] = i < n ? in[i] : INT64_MAX;
  return true;
}

bool BitonicStep(void** args) {   // (buf, npad, k, j)
  int64_t* buf = Arg<int64_t*>(args, 0);
  const int64_t npad = Arg<int64_t>(args, 1), k = Arg<int64_t>(args, 2), j = Arg<int64_t>(args, 3);
  for (int64_t i = 0; i < npad; ++i) {
    const int64_t l = i ^ j;
    if (l > i) {
      const bool up = (i & k) == 0;
      const int64_t a = buf[i], b = buf[l];
      if ((a > b) == up) { buf[i] = b; buf[l] = a; }
    }
  }
  return true;
}

bool CompactSorted(void** args) {   // (sorted, n, out, d_count)
  const int64_t* s = Arg<const int64_t*>(args, 0);
  const int64_t n = Arg<int64_t>(args, 1);
  int64_t* out = Arg<int64_t*>(args, 2);
  int64_t cnt = 0;
  for (int64_t i = 0; i < n; ++i)
    if (i == 0 || s[i] != s[i - 1]) out[cnt++] = s[i];
  *Arg<int64_t*>(args, 3) = cnt;
  return true;
}

bool Gather(void** args) {     // (table, ids, d_count, L, out_val, out_idx, vec)
  const float* table = Arg<const float*>(args, 0);
  const int64_t* ids = Arg<const int64_t*>(args, 1);
  const int64_t n = *Arg<const int64_t*>(args, 2);
  const int64_t L = Arg<int64_t>(args, 3);
  float* out_val = Arg<float*>(args, 4);
  int64_t* out_idx = Arg<int64_t*>(args, 5);
  for (int64_t j = 0; j < n; ++j) {
    const int64_t id = ids[j];
    if (out_idx != ids) out_idx[j] = id;
    std::memcpy(out_val + j * L, table + id * L, L * sizeof(float));
  }
  return true;
}

bool Scatter(void** args) {    // (table, idx, d_nnz, L, val)
  float* table = Arg<float*>(args, 0);
  const int64_t* idx = Arg<const int64_t*>(args, 1);
  const int64_t n = *Arg<const int64_t*>(args, 2);
  const int64_t L = Arg<int64_t>(args, 3);
  const float* val = Arg<const float*>(args, 4);
  for (int64_t j = 0; j < n; ++j) std::memcpy(table + idx[j] * L, val + j * L, L * sizeof(float));
  return true;
}

bool CastIds(const std::vector<std::string>& t, void** args) {   // rsp_cast_ids_kernel<T>(src, dst, n)
  const void* src = Arg<const void*>(args, 0);
  int64_t* dst = Arg<int64_t*>(args, 1);
  const int64_t n = Arg<int64_t>(args, 2);
  for (int64_t i = 0; i < n; ++i) {
    if (t[0] == "float") dst[i] = static_cast<int64_t>(static_cast<const float*>(src)[i]);
    else if (t[0] == "double") dst[i] = static_cast<int64_t>(static_cast<const double*>(src)[i]);
    else if (t[0] == "int") dst[i] = static_cast<const int32_t*>(src)[i];
    else if (t[0] == "long") dst[i] = static_cast<const int64_t*>(src)[i];
    else return false;
  }
  return true;
}

// grid barrier of device_utils.cuh for the one sequential "block" of an emulated kernel: locally a no-op; across
// ranks (one process per GPU on the simulator) the real epoch exchange through the peer-mapped pads
void GridBarrierCross(const SyncArgs& s) {
  static const bool mp = getenv("MXKV_SIM_MP") != nullptr;
  if (!mp || s.world <= 1) return;
  uint32_t* g = s.self + kSigGridOff;
  const uint32_t epoch = g[kSigGridEpoch] + 1;
  g[kSigGridEpoch] = epoch;
  for (int q = 0; q < s.world; ++q)
    __atomic_store_n(s.peers[q] + kSigGridOff + kSigGridFlags + s.rank, epoch, __ATOMIC_RELEASE);
  for (int q = 0; q < s.world; ++q) {
    const uint32_t* mine = g + kSigGridFlags + q;
    long spins = 0;
    const time_t t0 = time(nullptr);
    while (static_cast<int32_t>(__atomic_load_n(mine, __ATOMIC_ACQUIRE) - epoch) < 0) {
      if ((++spins & 255) == 0) {
        if (spins > 4096) usleep(100); else sched_yield();
        if (time(nullptr) - t0 > 600) { fprintf(stderr, "[mxkv sim] grid barrier timed out\n"); abort(); }
      }
    }
  }
}

template <int OPT>
void PushFusedT(const RspSources& S, const RspRowArgs& A, const RspStage& St, const SyncArgs& sync) {
  const int64_t L = A.row_len;
  if (St.publish) {
    if (St.src_nnz > 0) {
      std::memcpy(St.dst_idx, St.src_idx, St.src_nnz * sizeof(int64_t));
      std::memcpy(St.dst_val, St.src_val, St.src_nnz * L * sizeof(float));
    }
    *St.dst_nnz = St.src_nnz;
    GridBarrierCross(sync);
  }
  std::vector<int64_t> nnz(S.n);
  for (int t = 0; t < S.n; ++t) nnz[t] = St.nnz_by_value ? St.nnz_val[t] : *S.nnz[t];
  std::vector<const int64_t*> ids(S.n);
  for (int t = 0; t < S.n; ++t) {
    ids[t] = S.idx[t];
    if (St.localize) {
      int64_t* dst = St.lidx + static_cast<int64_t>(t) * St.lcap;
      if (nnz[t] > St.lcap) { fprintf(stderr, "[mxkv sim] id list exceeds the local workspace\n"); abort(); }
      if (nnz[t] > 0) std::memcpy(dst, S.idx[t], nnz[t] * sizeof(int64_t));
      ids[t] = dst;
    }
  }
  for (int s = 0; s < S.n; ++s) {
    for (int64_t r = 0; r < nnz[s]; ++r) {
      const int64_t id = ids[s][r];
      int64_t pos[kMaxSrc];
      bool earlier = false;
      for (int t = 0; t < S.n; ++t) {
        const int64_t p = (t == s) ? r : LowerBound(ids[t], nnz[t], id);
        pos[t] = (p < nnz[t] && ids[t][p] == id) ? p : -1;
        if (t < s && pos[t] >= 0) earlier = true;
      }
      if (earlier) continue;
      for (int64_t c = 0; c < L; ++c) {
        float acc = 0.f;
        for (int t = s; t < S.n; ++t)
          if (pos[t] >= 0) acc = __fadd_rn(acc, S.val[t][pos[t] * L + c]);
        float* w = A.table + id * L + c;
        *w = rsp_lazy_update<OPT>(acc, *w, id * L + c, A);
      }
    }
  }
  if (St.publish) GridBarrierCross(sync);
}

bool PushFused(const std::vector<std::string>& t, void** args) {   // rsp_push_fused_kernel<OPT, VEC>(S, A, St, sync)
  const RspSources& S = Arg<RspSources>(args, 0);
  const RspRowArgs& A = Arg<RspRowArgs>(args, 1);
  const RspStage& St = Arg<RspStage>(args, 2);
  const SyncArgs& sync = Arg<SyncArgs>(args, 3);
  switch (atoi(t[0].c_str())) {
    case OPT_SGD: PushFusedT<OPT_SGD>(S, A, St, sync); break;
    case OPT_SGD_MOM: PushFusedT<OPT_SGD_MOM>(S, A, St, sync); break;
    case OPT_ADAM: PushFusedT<OPT_ADAM>(S, A, St, sync); break;
    default: return false;
  }
  return true;
}

bool PullFused(void** args) {   // (table, in, n, out_idx, d_count, L, out_val, vec, sync)
  const float* table = Arg<const float*>(args, 0);
  const int64_t* in = Arg<const int64_t*>(args, 1);
  const int64_t n = Arg<int64_t>(args, 2);
  int64_t* out_idx = Arg<int64_t*>(args, 3);
  int64_t* d_count = Arg<int64_t*>(args, 4);
  const int64_t L = Arg<int64_t>(args, 5);
  float* out_val = Arg<float*>(args, 6);
  SortedUnique(std::vector<int64_t>(in, in + n), out_idx, d_count);
  for (int64_t j = 0; j < *d_count; ++j) std::memcpy(out_val + j * L, table + out_idx[j] * L, L * sizeof(float));
  return true;
}

}  // namespace

// hostemu_rsp.cc: csrc/rsp_kernels.cu compiled for the host
bool HostEmuRspLaunch(const std::string& kernel, const std::vector<std::string>& targs, void** args, unsigned gx, unsigned gy,
                      unsigned block, size_t smem);

bool DispatchRsp(const LaunchInfo& info, const std::string& base, const std::vector<std::string>& t, void** args) {
  const size_t p = base.rfind("::");
  const std::string k = p == std::string::npos ? base : base.substr(p + 2);
  // the kernels run from their OWN source unless MXKV_SIM_RSP=semantic asks for the independent emulators below
  static const bool semantic = [] { const char* v = getenv("MXKV_SIM_RSP"); return v != nullptr && std::string(v) == "semantic"; }();
  if (!semantic && k.compare(0, 4, "rsp_") == 0) {
    if (HostEmuRspLaunch(k, t, args, info.grid, info.grid_y, info.block, info.smem)) return true;
    fprintf(stderr, "[mxkv sim] rsp_kernels.cu from source: no entry for %s\n", info.name.c_str());
    abort();
  }
  if (k == "rsp_first_kernel") return First(args);
  if (k == "rsp_scan_kernel") return Scan(args);
  if (k == "rsp_rank_kernel") return Rank(args);
  if (k == "rsp_rows_kernel") return Rows(t, args);
  if (k == "rsp_push_fused_kernel") return PushFused(t, args);
  if (k == "rsp_pull_fused_kernel") return PullFused(args);
  if (k == "rsp_unique_kernel") return Unique(args);
  if (k == "rsp_pad_kernel") return Pad(args);
  if (k == "rsp_bitonic_step_kernel") return BitonicStep(args);
  if (k == "rsp_compact_sorted_kernel") return CompactSorted(args);
  if (k == "rsp_gather_kernel") return Gather(args);
  if (k == "rsp_scatter_kernel") return Scatter(args);
  if (k == "rsp_set_i64_kernel") { *Arg<int64_t*>(args, 0) = Arg<int64_t>(args, 1); return true; }
  if (k == "rsp_cast_ids_kernel") return CastIds(t, args);
  return false;
}

}  // namespace sim
