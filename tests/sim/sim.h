// sim.h -- interface between the simulated CUDA runtime (fake_cudart.cc) and the kernel emulators
// (sim_kernels.cc).  Test infrastructure only.
#pragma once
#include <string>
#include <vector>

namespace sim {

struct LaunchInfo {
  std::string name;        // demangled kernel name, e.g. "void mxkv::kv_dense_kernel<float, 2, false, true>(mxkv::DenseLaunch)"
  unsigned grid = 1, block = 1;
  unsigned grid_y = 1;
  size_t smem = 0;
  int device = 0;
};

// base name ("mxkv::kv_dense_kernel") and template arguments ("float", "2", "false", "true")
void ParseName(const std::string& demangled, std::string* base, std::vector<std::string>* targs);

// runs the emulator of the kernel; false if there is none
bool Dispatch(const LaunchInfo& info, void** args);

}  // namespace sim
