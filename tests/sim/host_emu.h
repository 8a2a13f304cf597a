// host_emu.h -- just enough of the CUDA execution model for g++ to compile a kernel's OWN source and run it on the
// CPU (test infrastructure; included by csrc/device_utils.cuh only when MXKV_HOST_EMU is defined, which only
// tests/sim/hostemu_tree.cc does).
//
// A block is a handful of OS threads (kHostEmuThreads, whatever block size the launch asked for: the kernels
// stride by blockDim.x), __syncthreads is a real barrier between them, __shared__ variables are statics (blocks
// run one after another), threadIdx / blockIdx / blockDim / gridDim are per-thread values.  16-byte accesses are
// aligned vector loads, so a packet the kernel addresses wrongly faults here as it would on the device.
// What this cannot show: anything about registers, occupancy, memory ordering between GPUs or speed.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#undef __device__
#undef __global__
#undef __host__
#undef __shared__
#undef __forceinline__
#undef __launch_bounds__
#define __device__
#define __global__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __launch_bounds__(...)

namespace hostemu {

constexpr int kHostEmuThreads = 8;

// a vector access the device would fault on (misaligned address) must not pass here either
template <typename P>
inline P* Aligned(P* p, uintptr_t bytes) {
  if (reinterpret_cast<uintptr_t>(p) & (bytes - 1)) {
    fprintf(stderr, "host_emu: misaligned %d-byte access at %p\n", static_cast<int>(bytes), static_cast<const void*>(p));
    std::abort();
  }
  return p;
}

struct Coords { uint3 thread_idx, block_idx; dim3 block_dim, grid_dim; };
inline Coords& Me() { static thread_local Coords c; return c; }

class BlockBarrier {
 public:
  explicit BlockBarrier(int n) : n_(n) {}
  void Wait() {
    std::unique_lock<std::mutex> lk(mu_);
    const unsigned gen = gen_;
    if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); return; }
    cv_.wait(lk, [&] { return gen_ != gen; });
  }
 private:
  std::mutex mu_;
  std::condition_variable cv_;
  int n_, count_ = 0;
  unsigned gen_ = 0;
};
inline BlockBarrier*& CurrentBarrier() { static thread_local BlockBarrier* b = nullptr; return b; }

// every block of the grid, one after another; its threads concurrently
template <typename Launch>
void RunGrid(void (*kernel)(Launch), const Launch& L, int grid, int /*threads_asked*/) {
  for (int b = 0; b < grid; ++b) {
    BlockBarrier bar(kHostEmuThreads);
    std::vector<std::thread> pool;
    for (int t = 0; t < kHostEmuThreads; ++t) {
      pool.emplace_back([&, t] {
        Coords& c = Me();
        c.thread_idx = make_uint3(t, 0, 0);
        c.block_idx = make_uint3(b, 0, 0);
        c.block_dim = dim3(kHostEmuThreads, 1, 1);
        c.grid_dim = dim3(grid, 1, 1);
        CurrentBarrier() = &bar;
        kernel(L);
      });
    }
    for (auto& th : pool) th.join();
  }
}

}  // namespace hostemu

#define threadIdx (::hostemu::Me().thread_idx)
#define blockIdx (::hostemu::Me().block_idx)
#define blockDim (::hostemu::Me().block_dim)
#define gridDim (::hostemu::Me().grid_dim)

inline void __syncthreads() { ::hostemu::CurrentBarrier()->Wait(); }
inline long long clock64() { return 0; }
inline void __trap() { std::abort(); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline double __dadd_rn(double a, double b) { return a + b; }
