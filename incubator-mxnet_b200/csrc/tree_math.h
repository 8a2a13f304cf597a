// tree_math.h -- one thread's evaluation of a reduction tree (MXNET_KVSTORE_USETREE=1, topology.h), in a header
// that compiles for the device (kernels.cu, the TREE instantiations) and for the host (tests/sim/sim_kernels.cc,
// tests/c/tree_host.cc): the same source the GPU executes is checked on CPU against the oracle's level-by-level
// restatement of CommDeviceTree::ReduceInner (oracle/oracle.py: sum_tree).
//
// The values of one element arrive in the tree's leaf order.  A program word says, after each value, how many of
// the pending partial sums are complete now (one 0 bit each, then a 1 bit): the running value absorbs them, newest
// first, and becomes the newest pending partial itself.  At most kTreeStack partials are pending (depth-3 trees);
// they live in registers, selected by comparisons -- no indexed (local-memory) access.
#pragma once
#include "kernels.h"

// host + device: MXKVB200TopologyRunProgram (c_api.cc, compiled by nvcc) runs the same code on the CPU
#if defined(__CUDACC__)
#define MXKV_TREE_HD __host__ __device__ __forceinline__
#else
#define MXKV_TREE_HD inline
#endif

namespace mxkv {

template <typename A, int N>
struct TreeSum {
  // the pending partials, oldest first.  Three separate arrays, not one [kTreeStack][N]: a select over rows of one
  // array is turned back into an indexed access by the compiler and the whole object then lives in local memory
  static_assert(kTreeStack == 3, "TreeSum spells its three slots out");
  A s0[N], s1[N], s2[N];
  int sp;
  uint32_t prog;

  MXKV_TREE_HD void begin(uint32_t program) {
    sp = 0;
    prog = program;
#if !defined(__CUDA_ARCH__)      // (host builds only: keeps -Wmaybe-uninitialized quiet; a slot is written before it is read)
    for (int i = 0; i < N; ++i) s0[i] = s1[i] = s2[i] = A();
#endif
  }

  // add(l, r): the rounded sum of two partials (the one ElementwiseSum of two arrays the reference runs)
  template <typename Add>
  MXKV_TREE_HD void take(const A (&x)[N], Add add) {
    A t[N];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < N; ++i) t[i] = x[i];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < kTreeStack; ++r) {
      if ((prog & 1u) == 0u) {           // uniform over the block: the program belongs to the work entry
        --sp;
        const bool is1 = sp == 1, is2 = sp == 2;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) {
          const A l = is2 ? s2[i] : (is1 ? s1[i] : s0[i]);
          t[i] = add(l, t[i]);
        }
        prog >>= 1;
      }
    }
    prog >>= 1;
    const bool to0 = sp == 0, to1 = sp == 1, to2 = sp == 2;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < N; ++i) {
      s0[i] = to0 ? t[i] : s0[i];
      s1[i] = to1 ? t[i] : s1[i];
      s2[i] = to2 ? t[i] : s2[i];
    }
    ++sp;
  }

  MXKV_TREE_HD void result(A (&out)[N]) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < N; ++i) out[i] = s0[i];
  }
};

// float32 partials: round to nearest, never contracted with a neighbouring multiply
struct TreeAddF32 {
  MXKV_TREE_HD float operator()(float l, float r) const {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(l, r);
#else
    return l + r;
#endif
  }
};

}  // namespace mxkv
