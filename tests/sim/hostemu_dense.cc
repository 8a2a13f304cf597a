// hostemu_dense.cc -- csrc/kernels.cu, the file the device executes, compiled by g++ and run on the CPU (host_emu.h):
// the per-thread kernel (kv_dense_kernel), the shared-memory staged kernel (kv_dense_bulk_kernel: its mbarrier /
// bulk-copy ring modelled as phases and byte counts) and the typed sum (kv_sum_typed_kernel).  The simulated runtime
// hands dense launches to THIS code, so the `-m gpu` parity tests of the CPU suite walk the kernels' own chunk and
// tile walks, descriptor staging, vector / scalar splits, summation orders, optimizer arithmetic and stores.
// (Test infrastructure.  MXKV_SIM_DENSE=semantic selects the independent emulators of sim_kernels.cc; the NVSwitch
// multicast kernel has no CPU model and is compiled out.)
#define MXKV_HOST_EMU 1
// (the packets read a uint4 as T[]: fine for nvcc, undefined for g++ unless built with -fno-strict-aliasing, which
// build_sim.py passes)
#pragma GCC diagnostic ignored "-Wmaybe-uninitialized"
#pragma GCC diagnostic ignored "-Wunused-function"
#define mxkv mxkv_hostemu
#include "../../incubator-mxnet_b200/csrc/kernels.cu"
#undef mxkv

namespace sim {

// `launch` points to a mxkv::DenseLaunch (same layout: same header, other namespace name)
bool HostEmuDenseLaunch(const void* launch) {
  return mxkv_hostemu::LaunchDenseHostEmu(*static_cast<const mxkv_hostemu::DenseLaunch*>(launch)) == 0;
}

}  // namespace sim
