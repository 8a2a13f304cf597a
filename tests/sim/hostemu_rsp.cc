// hostemu_rsp.cc -- csrc/rsp_kernels.cu (row_sparse union / gather-sum / lazy updates, the fused one-launch push and
// pull with their in-kernel grid barriers, unique, gather / scatter), the file the device executes, compiled by g++
// and run on the CPU with the launches' REAL block sizes (256 ... 1024 threads as user-level contexts: the kernels
// scan, vote and shuffle inside warps).  A launch keeps the product's grid, except the two fused kernels: their
// grid-wide barrier needs every block alive at once and `__shared__` is a static here, so they run as ONE block --
// they walk their work with grid strides, so one block does all of it, and the cross-GPU half of the barrier (one
// process per simulated GPU) is the real exchange through the peer-mapped pads.
// (Test infrastructure.  MXKV_SIM_RSP=semantic selects the independent emulators of sim_rsp.cc.)
#define MXKV_HOST_EMU 1
#pragma GCC diagnostic ignored "-Wmaybe-uninitialized"
#pragma GCC diagnostic ignored "-Wunused-function"
#define mxkv mxkv_hostemu
#include "gen/rsp_kernels.host.cu"
#undef mxkv
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace sim {

using namespace mxkv_hostemu;

namespace {

// cudaLaunchKernel's `args`: one pointer per kernel parameter, in order
template <typename... P, size_t... I>
void CallImpl(void (*k)(P...), dim3 grid, int threads, size_t smem, void** args, std::index_sequence<I...>) {
  ::hostemu::Launch(k, grid, threads, smem)(*static_cast<std::remove_reference_t<P>*>(args[I])...);
}
template <typename... P>
bool Call(void (*k)(P...), dim3 grid, int threads, size_t smem, void** args) {
  CallImpl(k, grid, threads, smem, args, std::index_sequence_for<P...>{});
  return true;
}

template <bool VEC>
bool Rows(int opt, dim3 g, int th, void** a) {
  switch (opt) {
    case OPT_NONE: return Call(rsp_rows_kernel<OPT_NONE, VEC>, g, th, 0, a);
    case OPT_SGD: return Call(rsp_rows_kernel<OPT_SGD, VEC>, g, th, 0, a);
    case OPT_SGD_MOM: return Call(rsp_rows_kernel<OPT_SGD_MOM, VEC>, g, th, 0, a);
    case OPT_ADAM: return Call(rsp_rows_kernel<OPT_ADAM, VEC>, g, th, 0, a);
    default: return false;
  }
}
template <bool VEC>
bool PushFused(int opt, int th, void** a) {
  const dim3 one(1, 1, 1);
  switch (opt) {
    case OPT_SGD: return Call(rsp_push_fused_kernel<OPT_SGD, VEC>, one, th, 0, a);
    case OPT_SGD_MOM: return Call(rsp_push_fused_kernel<OPT_SGD_MOM, VEC>, one, th, 0, a);
    case OPT_ADAM: return Call(rsp_push_fused_kernel<OPT_ADAM, VEC>, one, th, 0, a);
    default: return false;
  }
}

}  // namespace

// kernel: name without namespace; targs: its template arguments as the demangler prints them
bool HostEmuRspLaunch(const std::string& k, const std::vector<std::string>& t, void** a, unsigned gx, unsigned gy,
                      unsigned block, size_t smem) {
  const dim3 g(gx, gy, 1);
  const int th = static_cast<int>(block);
  auto flag = [&](size_t i) { return i < t.size() && (t[i] == "true" || t[i] == "1"); };
  if (k == "rsp_first_kernel") return Call(rsp_first_kernel, g, th, smem, a);
  if (k == "rsp_scan_kernel") return Call(rsp_scan_kernel, g, th, smem, a);
  if (k == "rsp_rank_kernel") return Call(rsp_rank_kernel, g, th, smem, a);
  if (k == "rsp_rows_kernel") return flag(1) ? Rows<true>(atoi(t[0].c_str()), g, th, a) : Rows<false>(atoi(t[0].c_str()), g, th, a);
  if (k == "rsp_push_fused_kernel") return flag(1) ? PushFused<true>(atoi(t[0].c_str()), th, a) : PushFused<false>(atoi(t[0].c_str()), th, a);
  if (k == "rsp_pull_fused_kernel") return Call(rsp_pull_fused_kernel, dim3(1, 1, 1), th, smem, a);
  if (k == "rsp_unique_kernel") return Call(rsp_unique_kernel, g, th, smem, a);
  if (k == "rsp_pad_kernel") return Call(rsp_pad_kernel, g, th, smem, a);
  if (k == "rsp_bitonic_step_kernel") return Call(rsp_bitonic_step_kernel, g, th, smem, a);
  if (k == "rsp_compact_sorted_kernel") return Call(rsp_compact_sorted_kernel, g, th, smem, a);
  if (k == "rsp_gather_kernel") return Call(rsp_gather_kernel, g, th, smem, a);
  if (k == "rsp_scatter_kernel") return Call(rsp_scatter_kernel, g, th, smem, a);
  if (k == "rsp_set_i64_kernel") return Call(rsp_set_i64_kernel, g, th, smem, a);
  if (k == "rsp_cast_ids_kernel") {
    if (t.empty()) return false;
    if (t[0] == "float") return Call(rsp_cast_ids_kernel<float>, g, th, smem, a);
    if (t[0] == "double") return Call(rsp_cast_ids_kernel<double>, g, th, smem, a);
    if (t[0] == "int") return Call(rsp_cast_ids_kernel<int32_t>, g, th, smem, a);
    if (t[0] == "long") return Call(rsp_cast_ids_kernel<int64_t>, g, th, smem, a);
  }
  return false;
}

}  // namespace sim
