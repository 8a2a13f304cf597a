// host_emu.h -- just enough of the CUDA execution model for g++ to compile a kernel's OWN source and run it on the
// CPU (test infrastructure; included by csrc/device_utils.cuh only when MXKV_HOST_EMU is defined, which only the
// tests/sim/hostemu_*.cc translation units do).
//
// A block is a handful of OS threads (kHostEmuThreads, whatever block size the launch asked for: the kernels
// stride by blockDim.x), __syncthreads is a real barrier between them, __shared__ variables are statics and the
// dynamic shared memory one buffer (blocks run one after another, launches are serialised), threadIdx / blockIdx /
// blockDim / gridDim are per-thread values.  16- and 8-byte accesses are checked for alignment, so a packet the
// kernel addresses wrongly aborts here as it would fault on the device.  The mbarrier / bulk-copy pair of the staged
// kernel is modelled as what the PTX says: expect_tx arms a phase with a byte count, every copy completes its bytes,
// the phase flips when the count is back to zero, waiters spin on the phase parity.
// What this cannot show: anything about registers, occupancy, asynchrony of the copy engine, memory ordering between
// GPUs, or speed.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
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

// a flag load of a spin loop: the peer it waits for may need this core (several ranks of 8 threads share the machine)
inline uint32_t PoliteLoad(const uint32_t* p, int order) {
  const uint32_t v = __atomic_load_n(p, order);
  std::this_thread::yield();
  return v;
}

struct Coords { uint3 thread_idx, block_idx; dim3 block_dim, grid_dim; };
inline Coords& Me() { static thread_local Coords c; return c; }

// sense-reversing barrier: a few spins, then yield (the suite runs several processes of these threads on few cores)
class BlockBarrier {
 public:
  explicit BlockBarrier(int n) : n_(n) {}
  void Wait() {
    const unsigned gen = gen_.load(std::memory_order_acquire);
    if (count_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
      count_.store(0, std::memory_order_relaxed);
      gen_.store(gen + 1, std::memory_order_release);
      return;
    }
    for (int spins = 0; gen_.load(std::memory_order_acquire) == gen; ++spins)
      if (spins > 32) std::this_thread::yield();
  }
 private:
  const int n_;
  std::atomic<int> count_{0};
  std::atomic<unsigned> gen_{0};
};

// kHostEmuThreads workers that live as long as the process (never joined: they sleep between blocks)
class Pool {
 public:
  static Pool& Get() { static Pool* p = new Pool; return *p; }
  BlockBarrier& barrier() { return barrier_; }
  std::mutex& launch_mutex() { return launch_mu_; }
  unsigned char* dynamic_smem(size_t bytes) {
    if (bytes + 128 > smem_.size()) smem_.resize(bytes + 128);
    return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_.data()) + 127) & ~uintptr_t(127));
  }
  unsigned char*& current_smem() { return current_smem_; }
  // fn(t) on every worker at once; returns when all are done
  void Run(const std::function<void(int)>& fn) {
    std::unique_lock<std::mutex> lk(mu_);
    job_ = &fn;
    remaining_ = kHostEmuThreads;
    ++generation_;
    cv_.notify_all();
    done_.wait(lk, [&] { return remaining_ == 0; });
    job_ = nullptr;
  }
 private:
  Pool() : barrier_(kHostEmuThreads) {
    for (int t = 0; t < kHostEmuThreads; ++t) {
      std::thread([this, t] {
        unsigned seen = 0;
        for (;;) {
          const std::function<void(int)>* job;
          {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return generation_ != seen; });
            seen = generation_;
            job = job_;
          }
          (*job)(t);
          std::unique_lock<std::mutex> lk(mu_);
          if (--remaining_ == 0) done_.notify_all();
        }
      }).detach();
    }
  }
  std::mutex mu_, launch_mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* job_ = nullptr;
  int remaining_ = 0;
  unsigned generation_ = 0;
  BlockBarrier barrier_;
  std::vector<unsigned char> smem_;
  unsigned char* current_smem_ = nullptr;
};

inline unsigned char* DynamicSmem() { return Pool::Get().current_smem(); }

// every block of the grid, one after another; its threads concurrently.  kernel(args...) is the __global__ function.
template <typename... P, typename... A>
void RunGridSmem(void (*kernel)(P...), int grid, size_t smem_bytes, const A&... args) {
  Pool& pool = Pool::Get();
  std::lock_guard<std::mutex> one_launch(pool.launch_mutex());
  pool.current_smem() = pool.dynamic_smem(smem_bytes);
  pool.Run([&](int t) {                 // one hand-over per launch; the blocks are separated by the block barrier
    Coords& c = Me();
    c.thread_idx = make_uint3(t, 0, 0);
    c.block_dim = dim3(kHostEmuThreads, 1, 1);
    c.grid_dim = dim3(grid, 1, 1);
    for (int b = 0; b < grid; ++b) {
      c.block_idx = make_uint3(b, 0, 0);
      kernel(args...);
      pool.barrier().Wait();            // nobody enters the next block (same statics) before all have left this one
    }
  });
}
template <typename Launch>
void RunGrid(void (*kernel)(Launch), const Launch& L, int grid, int /*threads_asked*/, size_t smem_bytes = 0) {
  RunGridSmem(kernel, grid, smem_bytes, L);
}

// ---- mbarrier with a transaction count + cp.async.bulk (staged kernel) ------------------------------------------------
// the 64-bit barrier word: bits 0..31 outstanding bytes of the armed phase, bit 32 armed, bits 33.. completed phases
inline std::mutex& MbarMutex() { static std::mutex m; return m; }
inline void MbarInit(uint64_t* bar) { std::lock_guard<std::mutex> lk(MbarMutex()); *bar = 0; }
inline void MbarComplete(uint64_t* bar) {      // (lock held) the phase flips once it is armed and nothing is outstanding
  if ((*bar & 0xFFFFFFFFull) == 0 && (*bar >> 32 & 1)) *bar = ((*bar >> 33) + 1) << 33;
}
inline void MbarExpectTx(uint64_t* bar, uint32_t bytes) {
  std::lock_guard<std::mutex> lk(MbarMutex());
  *bar = (*bar & ~0xFFFFFFFFull) | ((*bar & 0xFFFFFFFFull) + bytes) | (1ull << 32);
  MbarComplete(bar);
}
inline void BulkCopy(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  // cp.async.bulk: 16-byte aligned addresses, size a multiple of 16
  Aligned(dst, 16); Aligned(src, 16);
  if (bytes & 15) { fprintf(stderr, "host_emu: bulk copy of %u bytes (not a multiple of 16)\n", bytes); std::abort(); }
  std::memcpy(dst, src, bytes);
  std::lock_guard<std::mutex> lk(MbarMutex());
  if ((*bar & 0xFFFFFFFFull) < bytes) { fprintf(stderr, "host_emu: more bytes copied than the barrier expects\n"); std::abort(); }
  *bar -= bytes;
  MbarComplete(bar);
}
inline void MbarWait(const uint64_t* bar, uint32_t parity) {
  for (;;) {
    {
      std::lock_guard<std::mutex> lk(MbarMutex());
      if (((*bar >> 33) & 1) != parity) return;     // the phase with this parity has completed
    }
    std::this_thread::yield();
  }
}

}  // namespace hostemu

#define threadIdx (::hostemu::Me().thread_idx)
#define blockIdx (::hostemu::Me().block_idx)
#define blockDim (::hostemu::Me().block_dim)
#define gridDim (::hostemu::Me().grid_dim)

inline void __syncthreads() { ::hostemu::Pool::Get().barrier().Wait(); }
// ~SM cycles: the kernels' spin timeouts (MXKV_B200_SPIN_TIMEOUT_S at 1.9 GHz) then mean what they say
inline long long clock64() {
  return static_cast<long long>(std::chrono::duration_cast<std::chrono::nanoseconds>(
      std::chrono::steady_clock::now().time_since_epoch()).count() * 1.9);
}
inline void __trap() { std::abort(); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline double __dadd_rn(double a, double b) { return a + b; }
