// host_emu.h -- just enough of the CUDA execution model for g++ to compile a kernel's OWN source and run it on the
// CPU (test infrastructure; included by csrc/device_utils.cuh only when MXKV_HOST_EMU is defined, which only the
// tests/sim/hostemu_*.cc translation units do).
//
// A block is its real number of threads, run as user-level contexts on the calling OS thread and switched round-robin
// at barriers, warp exchanges and spin loops (FiberBlock) -- or, for kernels that only stride by blockDim.x, a handful
// of pooled OS threads that really run concurrently (MXKV_SIM_ENGINE=threads).  __syncthreads is a real barrier,
// __shared__ variables are statics and the dynamic shared memory one buffer (blocks run one after another, launches
// are serialised), threadIdx / blockIdx / blockDim / gridDim are per-context values.  16- and 8-byte accesses are
// checked for alignment, so a packet the kernel addresses wrongly aborts here as it would fault on the device.  Warp
// shuffles, ballots and block votes are exchanges between the 32 contexts of a warp.  The mbarrier / bulk-copy pair of
// the staged kernel is modelled as what the PTX says: expect_tx arms a phase with a byte count, every copy completes
// its bytes, the phase flips when the count is back to zero, waiters spin on the phase parity.
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
#include <string>
#include <thread>
#include <ucontext.h>
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
void FiberYieldIfAny();      // (defined with FiberBlock)
inline uint32_t PoliteLoad(const uint32_t* p, int order) {
  const uint32_t v = __atomic_load_n(p, order);
  FiberYieldIfAny();
  std::this_thread::yield();
  return v;
}

struct Coords { uint3 thread_idx, block_idx; dim3 block_dim, grid_dim; };
inline Coords& ThreadCoords() { static thread_local Coords c; return c; }

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
    Coords& c = ThreadCoords();
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
// (declared below) every block with its real thread count, as user-level contexts
template <typename... P, typename... A>
void RunGridFibers(void (*kernel)(P...), dim3 grid, int threads, size_t smem_bytes, const A&... args);

// Dense and tree kernels only stride by blockDim.x, so either engine runs them: the block's real thread count as
// user-level contexts on the calling thread (default: deterministic, no oversubscription when several test processes
// share the machine), or 8 pooled OS threads that really run concurrently (MXKV_SIM_ENGINE=threads).
template <typename Launch>
void RunGrid(void (*kernel)(Launch), const Launch& L, int grid, int threads_asked, size_t smem_bytes = 0) {
  static const bool os_threads = [] { const char* v = getenv("MXKV_SIM_ENGINE"); return v != nullptr && std::string(v) == "threads"; }();
  if (os_threads) { RunGridSmem(kernel, grid, smem_bytes, L); return; }
  if (threads_asked != 128 && threads_asked != 256 && threads_asked != 512 && threads_asked != 1024) threads_asked = 512;
  // these kernels never look at a neighbour's registers: 64 contexts walk the same indices as 512 at an eighth of the
  // switches (MXKV_SIM_REAL_THREADS=1: the launch's own block size)
  static const bool real = getenv("MXKV_SIM_REAL_THREADS") != nullptr;
  RunGridFibers(kernel, dim3(static_cast<unsigned>(grid), 1, 1), real ? threads_asked : 64, smem_bytes, L);
}

// ---- fibers: a block with its REAL number of threads, for kernels that talk inside warps ------------------------------
// The layer-wise-optimizer kernels reduce with __shfl_xor_sync over 32 lanes and add the warps' totals in warp order:
// their sums are only the hardware's sums when a block has the hardware's 512 (256) threads.  Those blocks run here as
// that many user-level contexts on the calling OS thread, switched round-robin at barriers and warp exchanges
// (deterministic; one block at a time, so `static` still stands in for __shared__).
// A context switch that saves what the SysV x86-64 ABI says a call preserves (rbx, rbp, r12-r15, the stack pointer)
// and nothing else: swapcontext also saves the signal mask, a system call per switch, and these kernels switch
// hundreds of thousands of times per launch.  Other architectures keep swapcontext.
#if defined(__x86_64__)
#define MXKV_FAST_FIBERS 1
extern "C" void mxkv_hostemu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.weak mxkv_hostemu_switch
.type mxkv_hostemu_switch,@function
mxkv_hostemu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size mxkv_hostemu_switch,.-mxkv_hostemu_switch
)");
#endif

class FiberBlock {
 public:
  static FiberBlock*& Current() { static FiberBlock* b = nullptr; return b; }
  struct Fiber {
#if defined(MXKV_FAST_FIBERS)
    void* sp = nullptr;
#else
    ucontext_t ctx;
#endif
    Coords coords;
    bool done = false;
  };
  FiberBlock(int threads) : fibers_(threads), warp_arrived_((threads + 31) / 32, 0), warp_gen_((threads + 31) / 32, 0),
                            slots_(2 * static_cast<size_t>(threads)), bank_of_(threads, 0) {}
  Fiber& cur() { return fibers_[cur_]; }
  void Yield() {
#if defined(MXKV_FAST_FIBERS)
    mxkv_hostemu_switch(&fibers_[cur_].sp, sched_sp_);
#else
    swapcontext(&fibers_[cur_].ctx, &sched_);
#endif
  }
  void SyncThreads() {
    const unsigned gen = gen_;
    if (++arrived_ == live_) { arrived_ = 0; ++gen_; return; }
    while (gen_ == gen) Yield();
  }
  // __syncthreads_and / _or: two banks chosen by the barrier generation; the bank of the NEXT generation is reset before
  // this barrier, while nobody can be writing it yet
  int SyncThreadsVote(int pred, bool is_and) {
    const unsigned g = gen_ & 1;
    vote_[g ^ 1][0] = 1; vote_[g ^ 1][1] = 0;
    if (!pred) vote_[g][0] = 0; else vote_[g][1] = 1;
    SyncThreads();
    return is_and ? vote_[g][0] : vote_[g][1];
  }
  // all lanes of the calling fiber's warp (the kernels only use full masks outside divergent code)
  void SyncWarp() {
    const int w = cur_ >> 5;
    const int lanes = std::min<int>(32, static_cast<int>(fibers_.size()) - w * 32);
    const unsigned gen = warp_gen_[w];
    if (++warp_arrived_[w] == lanes) { warp_arrived_[w] = 0; ++warp_gen_[w]; return; }
    while (warp_gen_[w] == gen) Yield();
  }
  // Two banks of slots, used alternately: a lane can only write the bank being read by a slower lane of its warp after
  // it has passed the NEXT exchange's barrier, which that slower lane must reach first -- one barrier per exchange.
  template <typename T>
  T Exchange(T v, int src_lane) {           // value of lane `src_lane` of my warp (own value when that lane does not exist)
    static_assert(sizeof(T) <= 8, "shuffles move up to 8 bytes");
    const int w = cur_ >> 5, lane = cur_ & 31;
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    uint64_t* bank = slots_.data() + (bank_of_[cur_] ^= 1) * fibers_.size();
    bank[cur_] = bits;
    SyncWarp();
    const int lanes = std::min<int>(32, static_cast<int>(fibers_.size()) - w * 32);
    const int src = (src_lane >= 0 && src_lane < lanes) ? src_lane : lane;
    T out;
    std::memcpy(&out, &bank[w * 32 + src], sizeof(T));
    return out;
  }
  unsigned Ballot(bool pred) {
    const int w = cur_ >> 5;
    uint64_t* bank = slots_.data() + (bank_of_[cur_] ^= 1) * fibers_.size();
    bank[cur_] = pred ? 1 : 0;
    SyncWarp();
    const int lanes = std::min<int>(32, static_cast<int>(fibers_.size()) - w * 32);
    unsigned m = 0;
    for (int l = 0; l < lanes; ++l) if (bank[w * 32 + l]) m |= 1u << l;
    return m;
  }
  template <typename F>
  void Run(dim3 block, dim3 grid, F&& body) {
    static std::vector<unsigned char> stacks;             // reused from launch to launch
    constexpr size_t kStack = 128 * 1024;
    if (stacks.size() < fibers_.size() * kStack) stacks.resize(fibers_.size() * kStack);
    body_ = [&body] { body(); };
    live_ = static_cast<int>(fibers_.size());
    for (size_t t = 0; t < fibers_.size(); ++t) {
      Fiber& f = fibers_[t];
      f.coords.thread_idx = make_uint3(static_cast<unsigned>(t), 0, 0);
      f.coords.block_idx = make_uint3(block.x, block.y, 0);
      f.coords.block_dim = dim3(static_cast<unsigned>(fibers_.size()), 1, 1);
      f.coords.grid_dim = grid;
      unsigned char* top = stacks.data() + (t + 1) * kStack;
#if defined(MXKV_FAST_FIBERS)
      // a fresh stack as mxkv_hostemu_switch expects to find one: six saved registers, then the address to `ret` to.
      // 16-byte alignment: at Entry's first instruction rsp % 16 == 8, as after a call
      void** sp = reinterpret_cast<void**>(reinterpret_cast<uintptr_t>(top) & ~uintptr_t(15));
      *--sp = nullptr;                                   // (keeps the alignment; Entry never returns)
      *--sp = reinterpret_cast<void*>(&FiberBlock::Entry);
      for (int r = 0; r < 6; ++r) *--sp = nullptr;
      f.sp = sp;
#else
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = top - kStack;
      f.ctx.uc_stack.ss_size = kStack;
      f.ctx.uc_link = &sched_;
      makecontext(&f.ctx, reinterpret_cast<void (*)()>(&FiberBlock::Entry), 0);
#endif
    }
    FiberBlock* prev = Current();
    Current() = this;
    int remaining = live_;
    while (remaining > 0) {
      for (cur_ = 0; cur_ < static_cast<int>(fibers_.size()); ++cur_) {
        if (fibers_[cur_].done) continue;
#if defined(MXKV_FAST_FIBERS)
        mxkv_hostemu_switch(&sched_sp_, fibers_[cur_].sp);
#else
        swapcontext(&sched_, &fibers_[cur_].ctx);
#endif
        if (fibers_[cur_].done) { --remaining; --live_; }
      }
    }
    Current() = prev;
  }
 private:
  static void Entry() {
    FiberBlock* b = Current();
    b->body_();
    b->fibers_[b->cur_].done = true;
#if defined(MXKV_FAST_FIBERS)
    for (;;) b->Yield();                    // back to the scheduler for good (a finished fiber is never resumed)
#endif                                      // (ucontext: uc_link takes the context back to the scheduler)
  }
  std::vector<Fiber> fibers_;
#if !defined(MXKV_FAST_FIBERS)
  ucontext_t sched_;
#endif
  std::function<void()> body_;
  int cur_ = 0, arrived_ = 0, live_ = 0;
  unsigned gen_ = 0;
  std::vector<int> warp_arrived_;
  std::vector<unsigned> warp_gen_;
  std::vector<uint64_t> slots_;
  std::vector<unsigned char> bank_of_;
  int vote_[2][2] = {{1, 0}, {1, 0}};       // [bank][and, or]
#if defined(MXKV_FAST_FIBERS)
  void* sched_sp_ = nullptr;
#endif
};

inline void FiberYieldIfAny() { if (FiberBlock* b = FiberBlock::Current()) b->Yield(); }

// every block of the grid (x fastest), one after another, each with `threads` fibers
template <typename... P, typename... A>
void RunGridFibers(void (*kernel)(P...), dim3 grid, int threads, size_t smem_bytes, const A&... args) {
  Pool& pool = Pool::Get();
  std::lock_guard<std::mutex> one_launch(pool.launch_mutex());
  pool.current_smem() = pool.dynamic_smem(smem_bytes);
  for (unsigned by = 0; by < grid.y; ++by) {
    for (unsigned bx = 0; bx < grid.x; ++bx) {
      FiberBlock block(threads);
      block.Run(dim3(bx, by, 1), grid, [&] { kernel(args...); });
    }
  }
}

// `kernel<<<grid, threads, smem, stream>>>(args)` of a .cu file compiled for the host: build_sim.py rewrites the launch
// into hostemu::Launch(kernel, grid, threads, smem)(args)
template <typename... P>
struct Launcher {
  void (*kernel)(P...);
  dim3 grid;
  int threads;
  size_t smem;
  template <typename... A>
  void operator()(const A&... args) const { RunGridFibers(kernel, grid, threads, smem, static_cast<P>(args)...); }
};
template <typename... P>
Launcher<P...> Launch(void (*kernel)(P...), dim3 grid, int threads, size_t smem) {
  return Launcher<P...>{kernel, grid, threads, smem};
}
template <typename... P>
Launcher<P...> Launch(void (*kernel)(P...), int64_t grid, int threads, size_t smem) {
  return Launcher<P...>{kernel, dim3(static_cast<unsigned>(grid), 1, 1), threads, smem};
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
    if (FiberBlock::Current() != nullptr) FiberYieldIfAny(); else std::this_thread::yield();
  }
}

}  // namespace hostemu

namespace hostemu {
inline Coords& Me() { FiberBlock* b = FiberBlock::Current(); return b != nullptr ? b->cur().coords : ThreadCoords(); }
}
#define threadIdx (::hostemu::Me().thread_idx)
#define blockIdx (::hostemu::Me().block_idx)
#define blockDim (::hostemu::Me().block_dim)
#define gridDim (::hostemu::Me().grid_dim)

inline void __syncthreads() {
  if (::hostemu::FiberBlock* b = ::hostemu::FiberBlock::Current()) b->SyncThreads();
  else ::hostemu::Pool::Get().barrier().Wait();
}
inline int __syncthreads_and(int pred) { return ::hostemu::FiberBlock::Current()->SyncThreadsVote(pred, true); }
inline int __syncthreads_or(int pred) { return ::hostemu::FiberBlock::Current()->SyncThreadsVote(pred, false); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz(static_cast<unsigned>(v)); }
// warp exchanges (fiber blocks only; every lane of the warp calls)
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  ::hostemu::FiberBlock* b = ::hostemu::FiberBlock::Current();
  return b->Exchange(v, static_cast<int>(threadIdx.x & 31) ^ lane_mask);
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
  ::hostemu::FiberBlock* b = ::hostemu::FiberBlock::Current();
  return b->Exchange(v, static_cast<int>(threadIdx.x & 31) - static_cast<int>(delta));
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src_lane) {
  return ::hostemu::FiberBlock::Current()->Exchange(v, src_lane & 31);
}
inline unsigned __ballot_sync(unsigned, int pred) { return ::hostemu::FiberBlock::Current()->Ballot(pred != 0); }
// occupancy queries of a launcher compiled for the host (the C++ overloads exist under nvcc only)
template <typename... P>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, void (*)(P...), int, size_t) { *n = 2; return cudaSuccess; }
template <typename... P>
inline cudaError_t cudaFuncSetAttribute(void (*)(P...), cudaFuncAttribute, int) { return cudaSuccess; }
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
template <typename T> inline T __ldg(const T* p) { return *p; }

// multimem.ld_reduce / multimem.st on an address of a multicast mapping (fake_driver.cc resolves it into the local
// mappings of every device's bound allocation): the sum of the n copies, added in device order -- the switch's own
// order is unspecified, results agree with any order to float32 rounding --, and a store into all of them
extern "C" int mxkv_sim_multimem(const void* p, void** out);
namespace hostemu {
inline float4 MultimemLoadReduce(const void* p) {
  void* c[8];
  const int n = mxkv_sim_multimem(Aligned(p, 16), c);
  float4 s = *static_cast<const float4*>(c[0]);
  for (int i = 1; i < n; ++i) {
    const float4 v = *static_cast<const float4*>(c[i]);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  return s;
}
inline void MultimemStore(void* p, const float4& v) {
  void* c[8];
  const int n = mxkv_sim_multimem(Aligned(p, 16), c);
  for (int i = 0; i < n; ++i) *static_cast<float4*>(c[i]) = v;
}
}  // namespace hostemu
