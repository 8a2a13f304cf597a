// norm_kernels.h -- layer-wise adaptive optimizers (LAMB, LANS, LARS) fused with the gradient
// exchange.  These optimizers need per-tensor L2 norms of the weight and of the update before a
// single element can be written, so the push becomes a short sequence of launches over the same
// work list instead of one:
//
//   first    : gather-sum the n gradient replicas (NVLink peer loads, reference association
//              order) and -- LAMB -- run step 1 (mean/var update, update direction), or --
//              LARS/LANS -- park the merged gradient; per-chunk partial sums of squares
//              (what the reference computes with multi_sum_sq, contrib/multi_sum_sq.cu:85-121)
//              and a non-finite count (all_finite.cu:33-66) fall out of the same pass
//   finalize : fixed-order sum of the chunk partials into this rank's per-key totals
//   mid      : LANS -- step 1 needs the gradient norm first (multi_lans-inl.h:352-362); also LAMB
//              when the update has to be skipped on overflow (nothing may be committed before
//              every gradient of the call is known to be finite, gluon/trainer.py:445-448)
//   apply    : every rank adds the ranks' totals in rank order (bit-identical trust ratios
//              everywhere), applies step 2 to its shard and stores the new weight to every
//              replica (the all-gather half of the two-shot exchange)
//
// Large keys are sharded across the ranks exactly like the plain optimizers (state and the
// temporaries of a key live shard-wise on their owner); small keys are reduced and updated
// redundantly by every rank, which then needs no cross-rank norm at all.
#pragma once
#include "kernels.h"

namespace mxkv {

enum NormKind : int {
  NORM_LAMB = 0,   // contrib/multi_lamb.cc:36-120 (step 1 / step 2), python/mxnet/optimizer/lamb.py
  NORM_LANS = 1,   // contrib/multi_lans.cc:36-130, python/mxnet/optimizer/lans.py
  NORM_LARS = 2    // python/mxnet/optimizer/lars.py:117-133 (_get_lars) + sgd_update / sgd_mom_update
};

// per-key totals, one float each (this rank's contribution; ranks are added in rank order)
constexpr int kNrmW = 0;        // sum w^2
constexpr int kNrmG = 1;        // LAMB: sum ghat^2; LARS/LANS: sum (rescale*g)^2
constexpr int kNrmM = 2;        // LANS: sum temp_m^2
constexpr int kNrmG2 = 3;       // LANS: sum temp_g^2
constexpr int kNrmBad = 4;      // number of non-finite elements of the merged gradient
constexpr int kNrmFloats = 8;
constexpr int kPsumStride = 3;  // floats per chunk in the per-launch partial buffer

struct alignas(16) NormWork {
  const void* src[kMaxSrc];
  void* out[kMaxOut];
  const void* w;       // stored weight, key dtype
  float* w32;          // fp32 master (multi precision)
  float* s0;           // LAMB/LANS mean | LARS momentum
  float* s1;           // LAMB/LANS var
  float* aux0;         // LAMB: update direction; LARS: merged gradient; LANS: merged gradient, then temp_g
  float* aux1;         // LANS: temp_m
  float* psum;         // this entry's chunk partials [chunks x kPsumStride]
  float* nrm;          // this rank's totals [kNrmFloats]
  const float* nrm_peer[kMaxRanks];   // every contributing rank's totals as mapped here
  int64_t begin, end;
  double lr_d;         // per-key learning rate before rounding (LARS multiplies in double, lars.py:258-260)
  int64_t reserved_;
  float lr, wd;
  float c1, c2;        // bias-correction denominators 1 - beta^t (host, float pow like the CPU kernel)
  int n_src, n_out;
  int flags;           // bit 0: vector path allowed; bit 1: no trust ratio for this key (LARS gamma/beta/bias)
  int norm_world;      // number of ranks whose totals are added (1: this rank holds the whole key)
};
static_assert(sizeof(NormWork) == 512, "NormWork layout");

struct NormLaunch {
  const NormWork* works;
  const int64_t* chunk_prefix;
  int nworks;
  int64_t total_chunks;
  int dtype;
  int kind;             // NormKind
  int multi_precision;
  int order;            // SumOrder
  int bias_correction;  // LAMB
  int has_momentum;     // LARS
  float rescale, clip, momentum, beta1, beta2, eps, lower_bound, upper_bound, lars_eta, lars_eps;
  int skip_nonfinite;   // 1: leave weight and state untouched when any merged gradient of the call is not finite
  int* overflow_flag;   // host-mapped word set to 1 by the apply kernel when that happened (may be null)
  // skip_nonfinite: addresses of the non-finite counts of EVERY key of the call (all launch classes,
  // all contributing ranks): the decision is one per push, identical on every rank
  const float* const* bad_list;
  int n_bad;
  SyncArgs sync;
  int grid;
  int chunk_elems;
};

int NormMaxGrid(int device);
// grad_only != 0: park the merged gradient in aux0 (LARS, LANS); 0: LAMB step 1
int LaunchNormFirst(const NormLaunch& L, int grad_only, cudaStream_t stream);
// psum[chunk][j] -> nrm[slot[j]] for j < nslots, one block per work entry
int LaunchNormFinalize(const NormLaunch& L, int nslots, int slot0, int slot1, int slot2, cudaStream_t stream);
// step 1 from the parked gradient: LANS always, LAMB when skip_nonfinite asks for the check first
int LaunchNormMid(const NormLaunch& L, cudaStream_t stream);
int LaunchNormApply(const NormLaunch& L, cudaStream_t stream);

// Stand-alone reductions over a list of arrays (the reference's multi_sum_sq / multi_all_finite
// operators): out[i] = sum(scale * x_i)^2, bad[i] = number of non-finite elements.
struct SumSqItem { const void* ptr; int64_t n; };
int LaunchMultiSumSq(const SumSqItem* d_items, const int64_t* d_chunk_prefix, int nitems, int64_t total_chunks,
                     int dtype, float scale, float* d_psum, float* d_out_sumsq, float* d_out_bad, int chunk_elems,
                     cudaStream_t stream);

// out[0] = (init ? 1 : out[0]), then 0 if any bad[i] > 0
int LaunchAllFiniteFlag(const float* d_bad, int n, float* d_out, int init, cudaStream_t stream);

}  // namespace mxkv
