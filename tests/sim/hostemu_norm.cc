// hostemu_norm.cc -- csrc/norm_kernels.cu (LAMB / LANS / LARS: first / finalize / mid / apply, multi_sum_sq,
// multi_all_finite), the file the device executes, compiled by g++ and run on the CPU.  These kernels reduce inside
// warps (__shfl_xor_sync) and add the warps' totals in warp order, so a block runs here with its REAL 512 (256)
// threads, as user-level contexts switched at barriers and warp exchanges (host_emu.h: FiberBlock): the sums of squares
// come out with the hardware's association.  The only rewrite is the launch syntax (build_sim.py: host_source).
// (Test infrastructure.  MXKV_SIM_NORM=semantic selects the independent emulators of sim_kernels.cc.)
#define MXKV_HOST_EMU 1
#pragma GCC diagnostic ignored "-Wmaybe-uninitialized"
#pragma GCC diagnostic ignored "-Wunused-function"
#define mxkv mxkv_hostemu
#include "gen/norm_kernels.host.cu"
#undef mxkv

namespace sim {

using namespace mxkv_hostemu;

// `launch` points to a mxkv::NormLaunch (same layout: same header, other namespace name); the launchers pick the
// instantiation from dtype / precision / kind exactly as the device launch does
bool HostEmuNormFirst(const void* launch, int grad_only) {
  return LaunchNormFirst(*static_cast<const NormLaunch*>(launch), grad_only, nullptr) == 0;
}
bool HostEmuNormMid(const void* launch) { return LaunchNormMid(*static_cast<const NormLaunch*>(launch), nullptr) == 0; }
bool HostEmuNormApply(const void* launch) { return LaunchNormApply(*static_cast<const NormLaunch*>(launch), nullptr) == 0; }
bool HostEmuNormFinalize(const void* works, const int64_t* prefix, int nworks, int nslots, int s0, int s1, int s2) {
  NormLaunch L{};
  L.works = static_cast<const NormWork*>(works);
  L.chunk_prefix = prefix;
  L.nworks = nworks;
  return LaunchNormFinalize(L, nslots, s0, s1, s2, nullptr) == 0;
}
void HostEmuSumSq(int type, const void* items, const int64_t* prefix, int nitems, int64_t total_chunks, float scale, float* psum,
                  int chunk_elems, int grid) {
  const SumSqItem* it = static_cast<const SumSqItem*>(items);
  switch (type) {
    case 0: ::hostemu::Launch(kv_sumsq_kernel<float>, grid, kNormThreads, 0)(it, prefix, nitems, total_chunks, scale, psum, chunk_elems); break;
    case 1: ::hostemu::Launch(kv_sumsq_kernel<__half>, grid, kNormThreads, 0)(it, prefix, nitems, total_chunks, scale, psum, chunk_elems); break;
    case 2: ::hostemu::Launch(kv_sumsq_kernel<__nv_bfloat16>, grid, kNormThreads, 0)(it, prefix, nitems, total_chunks, scale, psum, chunk_elems); break;
    default: ::hostemu::Launch(kv_sumsq_kernel<double>, grid, kNormThreads, 0)(it, prefix, nitems, total_chunks, scale, psum, chunk_elems); break;
  }
}
void HostEmuSumSqFinalize(const int64_t* prefix, const float* psum, float* out_sumsq, float* out_bad, int nitems) {
  ::hostemu::Launch(kv_sumsq_finalize_kernel, nitems, 256, 0)(prefix, psum, out_sumsq, out_bad);
}
void HostEmuAllFiniteFlag(const float* bad, int n, float* out, int init) {
  ::hostemu::Launch(kv_all_finite_flag_kernel, 1, 32, 0)(bad, n, out, init);
}

}  // namespace sim
