// hostemu_tree.cc -- csrc/tree_kernels.cu, the file the device executes, compiled by g++ and run on the CPU
// (host_emu.h): blocks with their real thread count as user-level contexts (or a few concurrent OS threads,
// MXKV_SIM_ENGINE=threads), a real barrier for __syncthreads, aligned 16-byte accesses.  The
// simulated runtime hands the tree-order launches to THIS code instead of a semantic emulator, so every tree test of
// the CPU suite walks the kernel's own chunk search, vector / scalar split, load batches, add schedule, optimizer
// and stores.  (Test infrastructure.  MXKV_SIM_TREE=semantic selects the independent emulator of sim_kernels.cc.)
#define MXKV_HOST_EMU 1
// (the packets read a uint4 as T[]: fine for nvcc, undefined for g++ unless built with -fno-strict-aliasing, which
// build_sim.py passes; without it the optimizer's weight loads came back as garbage at -O2)
#pragma GCC diagnostic ignored "-Wmaybe-uninitialized"
// this translation unit's copies of the engine's launchers and inline helpers must not collide with (or be
// interposed over) the product's own symbols in libmxkv_b200_sim.so
#define mxkv mxkv_hostemu
#include "../../incubator-mxnet_b200/csrc/tree_kernels.cu"
#undef mxkv

namespace sim {

// `launch` points to a mxkv::DenseLaunch (same layout: same header, other namespace name)
bool HostEmuTreeLaunch(const void* launch) {
  const mxkv_hostemu::DenseLaunch& L = *static_cast<const mxkv_hostemu::DenseLaunch*>(launch);
  return mxkv_hostemu::LaunchDenseTree(L, nullptr) == 0;
}

}  // namespace sim
