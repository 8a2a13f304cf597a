/*
 * kv_oracle.c -- CPU restatement of the MXNet KVStore gradient path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (incubator-mxnet_b200/csrc) never links or calls anything in here.
 *
 * Every function restates one reference routine in plain C and cites the
 * reference file:line (relative to /root/reference) it follows.  Arithmetic is
 * written exactly as the reference source spells it (separate multiply and add,
 * left-to-right association); the file MUST be compiled with
 * -ffp-contract=off so the compiler does not fuse a*b+c into an fma.
 *
 * The dense optimizer loops carry an OpenMP pragma (elementwise, order-free) like the
 * reference's CPU Kernel<OP, cpu>::Launch (src/operator/mxnet_op.h) so the CPU baseline can
 * use every host core.
 *
 * Pinning: the dense-sum routines are checked bit-for-bit against the
 * reference's own mshadow expression templates compiled from
 * /root/reference/3rdparty/mshadow (oracle/ref_harness.cc -> oracle/_ref/), and
 * against the known-answer tests of the reference test-suite re-expressed in
 * tests/ (see tests/golden/README.md).  The C++ optimizer operators cannot be built in
 * this container (libmxnet needs a BLAS and ~1.5k translation units), but the
 * reference's non-fused Python `step()` methods -- its own oracle for its fused kernels,
 * tests/python/unittest/test_optimizer.py -- can be executed: tests/golden/make_golden.py runs
 * SGD / Adam / AdamW / Test / LAMB / LANS / LARS `step` out of the reference tree and the optimizer
 * routines below are checked against those outputs (tests/golden/optimizer_steps.npz,
 * layerwise.npz) within that test's tolerances.  The sparse (lazy and standard) update kernels are
 * checked the same way against the classes the reference's tests hold them to -- PySparseSGD /
 * PySparseAdam of tests/python/unittest/test_optimizer.py and SGD.step on the densified gradient --
 * executed from the reference files (tests/golden/sparse_steps.npz).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef _Float16 f16_t;

/* ---- bf16 helpers (no reference: SURVEY "bf16 has no GPU path"; policy =
 * fp32 accumulate, round-to-nearest-even on store) ---- */
static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = ((uint32_t)h) << 16; float f; memcpy(&f, &u, 4); return f;
}
static inline uint16_t f32_to_bf16(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u); /* quiet NaN */
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

/* ===================================================================== */
/* Dense sums                                                             */
/* ===================================================================== */

/* CommDevice::Reduce -> ElementwiseSum (device order).
 * src/kvstore/comm.h:525-548 builds reduce = [merged(=src[0]), src[1], ...];
 * src/ndarray/ndarray_function-inl.h:457-486: n=2,3,4 -> out = in0+in1(+in2(+in3))
 * (C++ left-to-right), n>=5 -> out = in0; out += in_i for i=1.. .  Both are the
 * same association ((in0+in1)+in2)+...; every partial is rounded to DType. */
#define DEF_SUM_DEVICE(NAME, T)                                                  \
  void NAME(int n, const T* const* src, int64_t E, T* out) {                     \
    for (int64_t i = 0; i < E; ++i) {                                            \
      T acc = src[0][i];                                                         \
      for (int k = 1; k < n; ++k) acc = (T)(acc + src[k][i]);                    \
      out[i] = acc;                                                              \
    }                                                                            \
  }
DEF_SUM_DEVICE(kvo_sum_device_f32, float)
DEF_SUM_DEVICE(kvo_sum_device_f64, double)
DEF_SUM_DEVICE(kvo_sum_device_i32, int32_t)
DEF_SUM_DEVICE(kvo_sum_device_i64, int64_t)
DEF_SUM_DEVICE(kvo_sum_device_u8, uint8_t)
DEF_SUM_DEVICE(kvo_sum_device_i8, int8_t)

/* fp16: mshadow half_t operator+ computes float(a)+float(b) and rounds back to
 * half (3rdparty/mshadow/mshadow/half.h) -- i.e. NO fp32 accumulation across
 * inputs.  float add of two halves then RNE to half is the
 * correctly rounded half add (24 >= 2*11+2). */
void kvo_sum_device_f16(int n, const uint16_t* const* src, int64_t E, uint16_t* out) {
  for (int64_t i = 0; i < E; ++i) {
    f16_t acc; memcpy(&acc, &src[0][i], 2);
    for (int k = 1; k < n; ++k) {
      f16_t b; memcpy(&b, &src[k][i], 2);
      acc = (f16_t)((float)acc + (float)b);
    }
    memcpy(&out[i], &acc, 2);
  }
}

/* bf16: accumulate all n inputs in fp32 in the a10 order, one RNE at the end. */
void kvo_sum_device_bf16(int n, const uint16_t* const* src, int64_t E, uint16_t* out) {
  for (int64_t i = 0; i < E; ++i) {
    float acc = bf16_to_f32(src[0][i]);
    for (int k = 1; k < n; ++k) acc = acc + bf16_to_f32(src[k][i]);
    out[i] = f32_to_bf16(acc);
  }
}
/* same, result kept in fp32 (what the fused mp update consumes) */
void kvo_sum_device_bf16_f32out(int n, const uint16_t* const* src, int64_t E, float* out) {
  for (int64_t i = 0; i < E; ++i) {
    float acc = bf16_to_f32(src[0][i]);
    for (int k = 1; k < n; ++k) acc = acc + bf16_to_f32(src[k][i]);
    out[i] = acc;
  }
}
void kvo_sum_device_f16_f32out(int n, const uint16_t* const* src, int64_t E, float* out) {
  for (int64_t i = 0; i < E; ++i) {
    f16_t a; memcpy(&a, &src[0][i], 2);
    float acc = (float)a;
    for (int k = 1; k < n; ++k) { f16_t b; memcpy(&b, &src[k][i], 2); acc = acc + (float)b; }
    out[i] = acc;
  }
}

/* CommCPU::ReduceSumCPU, src/kvstore/comm.h:359-393: in place into dptr[0],
 * inputs taken in groups of up to four: in0 += ((in1+in2)+in3)+in4 .
 * (mshadow evaluates the right-hand expression per element, then +=.) */
static void sum_cpu_range_f32(int n, float* const* dptr, int64_t off, int64_t size) {
  float* in0 = dptr[0] + off;
  for (int i = 1; i < n; i += 4) {
    int left = n - i;
    const float* a = dptr[i] + off;
    if (left == 1) {
      for (int64_t j = 0; j < size; ++j) in0[j] = in0[j] + a[j];
    } else if (left == 2) {
      const float* b = dptr[i + 1] + off;
      for (int64_t j = 0; j < size; ++j) in0[j] = in0[j] + (a[j] + b[j]);
    } else if (left == 3) {
      const float* b = dptr[i + 1] + off; const float* c = dptr[i + 2] + off;
      for (int64_t j = 0; j < size; ++j) in0[j] = in0[j] + ((a[j] + b[j]) + c[j]);
    } else {
      const float* b = dptr[i + 1] + off; const float* c = dptr[i + 2] + off;
      const float* d = dptr[i + 3] + off;
      for (int64_t j = 0; j < size; ++j) in0[j] = in0[j] + (((a[j] + b[j]) + c[j]) + d[j]);
    }
  }
}

/* CommCPU::ReduceSumCPUImpl, comm.h:396-411: serial when total < bigarray_bound
 * (MXNET_KVSTORE_BIGARRAY_BOUND, default 1e6) or nthreads <= 1, else OpenMP
 * static schedule over 4096-element tasks with nthreads
 * (MXNET_KVSTORE_REDUCTION_NTHREADS, default 4). */
void kvo_sum_cpu_f32(int n, float* const* dptr, int64_t total, int nthreads, int64_t bigarray_bound) {
  int64_t step = bigarray_bound < 4096 ? bigarray_bound : 4096;
  if (step < 1) step = 1;
  int64_t ntask = (total + step - 1) / step;
  if (total < bigarray_bound || nthreads <= 1) {
    sum_cpu_range_f32(n, dptr, 0, total);
  } else {
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t j = 0; j < ntask; ++j) {
      int64_t b = j * step; if (b > total) b = total;
      int64_t e = (j + 1) * step; if (e > total) e = total;
      sum_cpu_range_f32(n, dptr, b, e - b);
    }
  }
}

/* ParallelCopy, src/common/utils.h:756-771 (pull on the CPU store): OMP loop
 * when size >= MXNET_CPU_PARALLEL_SIZE (200000). */
void kvo_parallel_copy_f32(float* dst, const float* src, int64_t size, int nthreads) {
  if (size >= 200000 && nthreads > 1) {
#pragma omp parallel for num_threads(nthreads)
    for (int64_t i = 0; i < size; ++i) dst[i] = src[i];
  } else {
    memcpy(dst, src, (size_t)size * sizeof(float));
  }
}

/* ===================================================================== */
/* Fused optimizer update kernels (dense)                                 */
/* ===================================================================== */
/* mshadow_op::clip, src/operator/mshadow_op.h:999-1009 */
static inline float clipf(float x, float b) { return x > b ? b : (x < -b ? -b : x); }

/* SGDKernel, src/operator/optimizer_op-inl.h:377-390 (DType = float) */
void kvo_sgd_update_f32(int64_t E, float* out, const float* w, const float* g,
                        float lr, float wd, float rescale, float clip) {
#pragma omp parallel for schedule(static) if (E >= 200000)
  for (int64_t i = 0; i < E; ++i) {
    float r = rescale * g[i];
    if (clip >= 0.0f) r = clipf(r, clip);
    r += wd * w[i];
    out[i] = w[i] - (lr * r);
  }
}

/* SGDMomKernel, optimizer_op-inl.h:590-606 */
void kvo_sgd_mom_update_f32(int64_t E, float* out, float* mom, const float* w, const float* g,
                            float lr, float wd, float momentum, float rescale, float clip) {
#pragma omp parallel for schedule(static) if (E >= 200000)
  for (int64_t i = 0; i < E; ++i) {
    float r = rescale * g[i];
    if (clip >= 0.0f) r = clipf(r, clip);
    r += wd * w[i];
    float m = mom[i];
    m *= momentum;
    m -= lr * r;
    mom[i] = m;
    out[i] = w[i] + m;
  }
}

/* MP_SGDKernel (optimizer_op-inl.h:642-658) / MP_SGDMomKernel (:681-701) on an
 * fp32 gradient.  The fused engine hands the update the fp32 sum of the
 * low-precision gradients; the low-precision weight copy is out_lp = (DType)w.
 * lp_kind: 0 = none, 1 = fp16, 2 = bf16. */
static inline void store_lp(uint16_t* out_lp, int64_t i, float w, int lp_kind) {
  if (lp_kind == 1) { f16_t h = (f16_t)w; memcpy(&out_lp[i], &h, 2); }
  else if (lp_kind == 2) out_lp[i] = f32_to_bf16(w);
}
void kvo_mp_sgd_update(int64_t E, uint16_t* out_lp, int lp_kind, float* w32, const float* g,
                       float lr, float wd, float rescale, float clip) {
#pragma omp parallel for schedule(static) if (E >= 200000)
  for (int64_t i = 0; i < E; ++i) {
    float w = w32[i];
    float r = rescale * g[i];
    if (clip >= 0.0f) r = clipf(r, clip);
    r += wd * w;
    w -= lr * r;
    w32[i] = w;
    store_lp(out_lp, i, w, lp_kind);
  }
}
void kvo_mp_sgd_mom_update(int64_t E, uint16_t* out_lp, int lp_kind, float* w32, float* mom,
                           const float* g, float lr, float wd, float momentum, float rescale,
                           float clip) {
#pragma omp parallel for schedule(static) if (E >= 200000)
  for (int64_t i = 0; i < E; ++i) {
    float w = w32[i];
    float m = mom[i];
    float r = rescale * g[i];
    if (clip >= 0.0f) r = clipf(r, clip);
    r += wd * w;
    m *= momentum;
    m -= lr * r;
    mom[i] = m;
    w = w + m;
    w32[i] = w;
    store_lp(out_lp, i, w, lp_kind);
  }
}

/* AdamUpdateKernel, optimizer_op-inl.h:1246-1269.  lr is already bias-corrected
 * on the host (python/mxnet/optimizer/adam.py:172-175, kvo_adam_lr below). */
void kvo_adam_update_f32(int64_t E, float* out, float* mean, float* var, const float* w,
                         const float* g, float lr, float wd, float beta1, float beta2,
                         float eps, float rescale, float clip) {
#pragma omp parallel for schedule(static) if (E >= 200000)
  for (int64_t i = 0; i < E; ++i) {
    float r = g[i] * rescale;
    if (clip >= 0.f) r = clipf(r, clip);
    r += w[i] * wd;
    float m = beta1 * mean[i] + (1.f - beta1) * r;
    float v = beta2 * var[i] + (1.f - beta2) * r * r;
    mean[i] = m; var[i] = v;
    out[i] = w[i] - lr * m / (sqrtf(v) + eps);
  }
}
/* same arithmetic on an fp32 master with a low-precision copy-out (engine
 * extension for bf16/fp16 weights; the reference has no mp_adam_update, only
 * mp_adamw -- SURVEY a26/a27). */
void kvo_mp_adam_update(int64_t E, uint16_t* out_lp, int lp_kind, float* w32, float* mean,
                        float* var, const float* g, float lr, float wd, float beta1, float beta2,
                        float eps, float rescale, float clip) {
#pragma omp parallel for schedule(static) if (E >= 200000)
  for (int64_t i = 0; i < E; ++i) {
    float w = w32[i];
    float r = g[i] * rescale;
    if (clip >= 0.f) r = clipf(r, clip);
    r += w * wd;
    float m = beta1 * mean[i] + (1.f - beta1) * r;
    float v = beta2 * var[i] + (1.f - beta2) * r * r;
    mean[i] = m; var[i] = v;
    w = w - lr * m / (sqrtf(v) + eps);
    w32[i] = w;
    store_lp(out_lp, i, w, lp_kind);
  }
}
/* host-side bias correction, adam.py:166-175, in double like Python */
double kvo_adam_lr(double lr, double beta1, double beta2, int t) {
  double coef1 = 1. - pow(beta1, (double)t);
  double coef2 = 1. - pow(beta2, (double)t);
  return lr * (sqrt(coef2) / coef1);
}

/* MPAdamWKernel, src/operator/contrib/adamw-inl.h:101-124 (fp32 master, no wd in
 * the gradient, decoupled decay).  lp_kind 0 -> plain fp32 weights: the
 * reference's fp32 AdamWUpdate (:165-192) is the same formula written as an
 * mshadow expression. */
void kvo_mp_adamw_update(int64_t E, uint16_t* out_lp, int lp_kind, float* w32, float* mean,
                         float* var, const float* g, float lr, float eta, float wd, float beta1,
                         float beta2, float eps, float rescale, float clip) {
#pragma omp parallel for schedule(static) if (E >= 200000)
  for (int64_t i = 0; i < E; ++i) {
    float w = w32[i];
    float sg = rescale * g[i];
    if (clip >= 0.0f) sg = clipf(sg, clip);
    float m = beta1 * mean[i] + (1.0f - beta1) * sg;
    float v = beta2 * var[i] + (1.0f - beta2) * (sg * sg);
    mean[i] = m; var[i] = v;
    w -= eta * (lr * m / (sqrtf(v) + eps) + wd * w);
    w32[i] = w;
    store_lp(out_lp, i, w, lp_kind);
  }
}

/* Test optimizer, python/mxnet/optimizer/optimizer.py:570-577:
 *   grad = rescale_grad * grad; weight[:] -= lr * (grad + wd * weight)
 * (NDArray ops: each is a separate fp32 elementwise op.) */
void kvo_test_update_f32(int64_t E, float* w, const float* g, float lr, float wd, float rescale) {
#pragma omp parallel for schedule(static) if (E >= 200000)
  for (int64_t i = 0; i < E; ++i) {
    float gr = rescale * g[i];
    float t = wd * w[i];
    float s = gr + t;
    float u = lr * s;
    w[i] = w[i] - u;
  }
}

/* ===================================================================== */
/* Row-sparse                                                              */
/* ===================================================================== */
static int cmp_i64(const void* a, const void* b) {
  int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
  return (x > y) - (x < y);
}
/* UniqueImpl<cpu>, src/kvstore/kvstore_utils.cc:32-44: sort + std::unique in
 * place; returns the number of unique values. */
int64_t kvo_unique_i64(int64_t* data, int64_t n) {
  if (n == 0) return 0;
  qsort(data, (size_t)n, sizeof(int64_t), cmp_i64);
  int64_t k = 1;
  for (int64_t i = 1; i < n; ++i) if (data[i] != data[k - 1]) data[k++] = data[i];
  return k;
}

/* ElementwiseSumRspImpl (GPU), src/ndarray/ndarray_function.cu:104-190:
 * out_idx = sorted union of the input row ids; out_val is zero-filled and each
 * input is added in input order (kernel ndarray_function-inl.cuh:34-62).
 * idx[k] has nnz[k] sorted unique int64 ids, val[k] is [nnz[k] x L].
 * out_idx/out_val must hold sum(nnz) rows; returns the union size. */
int64_t kvo_rsp_sum_f32(int n, const int64_t* const* idx, const float* const* val,
                        const int64_t* nnz, int64_t L, int64_t* out_idx, float* out_val) {
  int64_t tot = 0;
  for (int k = 0; k < n; ++k) { memcpy(out_idx + tot, idx[k], (size_t)nnz[k] * 8); tot += nnz[k]; }
  int64_t nu = kvo_unique_i64(out_idx, tot);
  memset(out_val, 0, (size_t)(nu * L) * sizeof(float));
  for (int k = 0; k < n; ++k) {
    int64_t p = 0;
    for (int64_t r = 0; r < nnz[k]; ++r) {
      int64_t id = idx[k][r];
      while (out_idx[p] != id) ++p;            /* both sorted */
      float* o = out_val + p * L; const float* v = val[k] + r * L;
      for (int64_t j = 0; j < L; ++j) o[j] = o[j] + v[j];
    }
  }
  return nu;
}

/* SparseRetainOpForwardRspImpl, src/operator/tensor/sparse_retain-inl.h:262-322:
 * out is zero-filled with idx.Size() rows, out_idx = idx; rows present in the
 * source are copied.  Source given as (src_idx sorted unique [src_nnz],
 * src_val [src_nnz x L]); src_nnz == num_rows is the "dense rsp" fast path. */
void kvo_sparse_retain_f32(const int64_t* src_idx, const float* src_val, int64_t src_nnz,
                           int64_t L, const int64_t* idx, int64_t nidx,
                           int64_t* out_idx, float* out_val) {
  memset(out_val, 0, (size_t)(nidx * L) * sizeof(float));
  for (int64_t i = 0; i < nidx; ++i) {
    out_idx[i] = idx[i];
    int64_t lo = 0, hi = src_nnz - 1, pos = -1;
    while (lo <= hi) {
      int64_t mid = lo + (hi - lo) / 2;
      if (src_idx[mid] == idx[i]) { pos = mid; break; }
      if (src_idx[mid] < idx[i]) lo = mid + 1; else hi = mid - 1;
    }
    if (pos >= 0) memcpy(out_val + i * L, src_val + pos * L, (size_t)L * sizeof(float));
  }
}

/* SGDDnsRspKernel (lazy update: only rows present in grad are touched),
 * optimizer_op-inl.h:414-465; dense weight [num_rows x L]. */
void kvo_sgd_rsp_lazy_f32(float* w, int64_t L, const int64_t* gidx, const float* gval, int64_t nnz,
                          float lr, float wd, float rescale, float clip) {
  for (int64_t r = 0; r < nnz; ++r) {
    float* wr = w + gidx[r] * L; const float* gr = gval + r * L;
    for (int64_t j = 0; j < L; ++j) {
      float x = rescale * gr[j];
      if (clip >= 0.0f) x = clipf(x, clip);
      x += wd * wr[j];
      wr[j] = wr[j] - (lr * x);
    }
  }
}
/* SGDMomDnsRspDnsKernel (lazy), optimizer_op-inl.h:724-790: same as SGDMomKernel
 * on the touched rows. */
void kvo_sgd_mom_rsp_lazy_f32(float* w, float* mom, int64_t L, const int64_t* gidx,
                              const float* gval, int64_t nnz, float lr, float wd, float momentum,
                              float rescale, float clip) {
  for (int64_t r = 0; r < nnz; ++r) {
    float* wr = w + gidx[r] * L; float* mr = mom + gidx[r] * L; const float* gr = gval + r * L;
    for (int64_t j = 0; j < L; ++j) {
      float x = rescale * gr[j];
      if (clip >= 0.0f) x = clipf(x, clip);
      x += wd * wr[j];
      float m = mr[j];
      m *= momentum;
      m -= lr * x;
      mr[j] = m;
      wr[j] = wr[j] + m;
    }
  }
}
/* AdamDnsRspDnsKernel (lazy), optimizer_op-inl.h:1305-1360 */
void kvo_adam_rsp_lazy_f32(float* w, float* mean, float* var, int64_t L, const int64_t* gidx,
                           const float* gval, int64_t nnz, float lr, float wd, float beta1,
                           float beta2, float eps, float rescale, float clip) {
  for (int64_t r = 0; r < nnz; ++r) {
    int64_t off = gidx[r] * L; const float* gr = gval + r * L;
    for (int64_t j = 0; j < L; ++j) {
      float x = gr[j] * rescale;
      if (clip >= 0.0f) x = clipf(x, clip);
      x += w[off + j] * wd;
      float m = beta1 * mean[off + j] + (1.f - beta1) * x;
      float v = beta2 * var[off + j] + (1.f - beta2) * x * x;
      mean[off + j] = m; var[off + j] = v;
      w[off + j] = w[off + j] - lr * m / (sqrtf(v) + eps);
    }
  }
}

/* Standard (lazy_update=false) updates of a dense weight with a row_sparse gradient.
 * SGD without momentum, SGDUpdateDnsRspImpl optimizer_op-inl.h:471-515: every row is first scaled by
 * (1 - lr*wd) (float arithmetic), then the rows of the gradient get w -= lr*clip(rescale*g) with wd = 0. */
void kvo_sgd_std_rsp_f32(float* w, int64_t R, int64_t L, const int64_t* gidx, const float* gval,
                         int64_t nnz, float lr, float wd, float rescale, float clip) {
  const float c = 1 - lr * wd;
  for (int64_t i = 0; i < R * L; ++i) w[i] = w[i] * c;
  for (int64_t r = 0; r < nnz; ++r) {
    float* wr = w + gidx[r] * L; const float* gr = gval + r * L;
    for (int64_t j = 0; j < L; ++j) {
      float x = rescale * gr[j];
      if (clip >= 0.0f) x = clipf(x, clip);
      x += 0.0f * wr[j];
      wr[j] = wr[j] - (lr * x);
    }
  }
}
/* AdamStdDnsRspDnsKernel, optimizer_op.cu:125-152, on the densified gradient (absent rows = 0):
 * as AdamUpdateKernel but the second moment uses (1-beta2)*square(g). */
void kvo_adam_std_update_f32(int64_t E, float* w, float* mean, float* var, const float* g, float lr,
                             float wd, float beta1, float beta2, float eps, float rescale, float clip) {
  for (int64_t i = 0; i < E; ++i) {
    float r = g[i] * rescale;
    if (clip >= 0.f) r = clipf(r, clip);
    r += w[i] * wd;
    float m = beta1 * mean[i] + (1.f - beta1) * r;
    float v = beta2 * var[i] + (1.f - beta2) * (r * r);
    mean[i] = m; var[i] = v;
    w[i] = w[i] - lr * m / (sqrtf(v) + eps);
  }
}

/* ===================================================================== */
/* Gradient compression (src/kvstore/gradient_compression-inl.h:44-227)    */
/* ===================================================================== */
/* quantize_2bit (:141-182): one thread per output byte = 4 values; codes
 * 11 (>= +thr), 10 (<= -thr), 00; residual carries the remainder. */
void kvo_quantize_2bit(int64_t E, const float* grad, float* residual, uint8_t* out, float thr) {
  int64_t nbytes = ((E + 15) / 16) * 4;
  const uint8_t posbits[] = {0xc0, 0x30, 0x0c, 0x03};
  const uint8_t negbits[] = {0x80, 0x20, 0x08, 0x02};
  for (int64_t b = 0; b < nbytes; ++b) {
    uint8_t c = 0;
    int64_t start = b << 2, end = start + 4 <= E ? start + 4 : E;
    for (int64_t i = start; i < end; ++i) {
      residual[i] += grad[i];
      if (residual[i] >= thr) { c |= posbits[i & 3]; residual[i] -= thr; }
      else if (residual[i] <= -thr) { c |= negbits[i & 3]; residual[i] -= -thr; }
    }
    out[b] = c;
  }
}
/* dequantize_2bit (:200-227) */
void kvo_dequantize_2bit(int64_t E, const uint8_t* in, float* out, float thr) {
  const uint8_t posbits[] = {0xc0, 0x30, 0x0c, 0x03};
  const uint8_t negbits[] = {0x80, 0x20, 0x08, 0x02};
  for (int64_t i = 0; i < E; ++i) {
    const uint8_t* ch = in + (i >> 4) * 4 + ((i & 15) >> 2);
    uint8_t masked = *ch & posbits[i & 3];
    if (masked == posbits[i & 3]) out[i] = thr;
    else if (masked == negbits[i & 3]) out[i] = -thr;
    else out[i] = 0;
  }
}
/* quantize_1bit (:44-80): one thread per output byte = 8 values; bit set when
 * residual > threshold (dequantised to +1), else -1. */
void kvo_quantize_1bit(int64_t E, const float* grad, float* residual, uint8_t* out, float thr) {
  int64_t nbytes = ((E + 31) / 32) * 4;
  const uint8_t bits[] = {0x80, 0x40, 0x20, 0x10, 0x08, 0x04, 0x02, 0x01};
  for (int64_t b = 0; b < nbytes; ++b) {
    uint8_t c = 0;
    int64_t start = b << 3, end = start + 8 <= E ? start + 8 : E;
    for (int64_t i = start; i < end; ++i) {
      residual[i] += grad[i];
      if (residual[i] > thr) { c |= bits[i & 7]; residual[i] -= 1; }
      else residual[i] += 1;
    }
    out[b] = c;
  }
}
/* dequantize_1bit (:97-123) */
void kvo_dequantize_1bit(int64_t E, const uint8_t* in, float* out, float thr) {
  (void)thr;
  const uint8_t bits[] = {0x80, 0x40, 0x20, 0x10, 0x08, 0x04, 0x02, 0x01};
  for (int64_t i = 0; i < E; ++i) {
    const uint8_t* ch = in + (i >> 5) * 4 + ((i & 31) >> 3);
    out[i] = ((*ch & bits[i & 7]) == bits[i & 7]) ? 1.0f : -1.0f;
  }
}

void kvo_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int kvo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Layer-wise adaptive optimizers.  CPU kernels of the reference:
 *   multi_sum_sq   src/operator/contrib/multi_sum_sq.cc:42-62 (CalcSumSq: sequential float sum)
 *   multi_lamb     src/operator/contrib/multi_lamb.cc:36-120 (step 1 / step 2)
 *   multi_lans     src/operator/contrib/multi_lans.cc:36-130
 *   LARS           python/mxnet/optimizer/lars.py:117-133 (_get_lars on float32 NDArrays)
 * The element arithmetic is restated operation by operation; the sums of squares are an input of
 * step 2 so that a test can feed the sequential sum (what the reference's CPU operator computes),
 * or a double-precision sum (the value every float summation order approximates).
 * ------------------------------------------------------------------------------------------ */
float kvo_sum_sq_f32(int64_t E, const float* x, float scale) {
  float sum = 0.f;
  for (int64_t j = 0; j < E; ++j) {
    float val = x[j];
    if (scale != 1.0f) val *= scale;
    sum += val * val;
  }
  return sum;
}
double kvo_sum_sq_f64(int64_t E, const float* x, float scale) {
  double sum = 0.0;
  for (int64_t j = 0; j < E; ++j) {
    float val = x[j];
    if (scale != 1.0f) val *= scale;
    sum += (double)val * (double)val;
  }
  return sum;
}
int64_t kvo_count_nonfinite_f32(int64_t E, const float* x) {
  int64_t bad = 0;
  for (int64_t j = 0; j < E; ++j) bad += !isfinite(x[j]);
  return bad;
}

/* MultiLAMBKernelStep1, multi_lamb.cc:36-82.  w is the fp32 weight (the master in the mp variant) */
void kvo_lamb_step1_f32(int64_t E, const float* w, const float* g, float* mean, float* var, float* temp_g,
                        float beta1, float beta2, float eps, float wd, float rescale, float clip,
                        int bias_correction, int step_count) {
  const float c1 = 1.0f - powf(beta1, (float)step_count);
  const float c2 = 1.0f - powf(beta2, (float)step_count);
#pragma omp parallel for schedule(static) if (E >= 200000)
  for (int64_t i = 0; i < E; ++i) {
    float scaled_grad = g[i] * rescale;
    if (clip >= 0.0f) scaled_grad = clipf(scaled_grad, clip);
    float m = beta1 * mean[i] + (1.0f - beta1) * scaled_grad;
    float v = beta2 * var[i] + (1.0f - beta2) * scaled_grad * scaled_grad;
    mean[i] = m; var[i] = v;
    float out;
    if (bias_correction) {
      float mean_hat = m / c1;
      float var_hat = v / c2;
      out = mean_hat / (sqrtf(var_hat) + eps) + wd * w[i];
    } else {
      out = m / (sqrtf(v) + eps) + wd * w[i];
    }
    temp_g[i] = out;
  }
}
/* MultiLAMBKernelStep2, multi_lamb.cc:84-120 */
void kvo_lamb_step2_f32(int64_t E, float* w, const float* temp_g, float lr, float sum_sq_w, float sum_sq_g,
                        float lower_bound, float upper_bound) {
  float r1 = sqrtf(sum_sq_w);
  float r2 = sqrtf(sum_sq_g);
  if (lower_bound >= 0) r1 = r1 > lower_bound ? r1 : lower_bound;
  if (upper_bound >= 0) r1 = r1 < upper_bound ? r1 : upper_bound;
  float r = (r1 == 0.0f || r2 == 0.0f) ? 1.0f : r1 / r2;
  float lr_adjusted = lr * r;
#pragma omp parallel for schedule(static) if (E >= 200000)
  for (int64_t i = 0; i < E; ++i) w[i] = w[i] - lr_adjusted * temp_g[i];
}

/* MultiLANSKernelStep1, multi_lans.cc:36-84 */
void kvo_lans_step1_f32(int64_t E, const float* w, const float* g, float* mean, float* var, float* temp_m,
                        float* temp_g, float beta1, float beta2, float eps, float wd, float rescale, float clip,
                        int step_count, float g_sq_norm) {
  const float c1 = 1.0f - powf(beta1, (float)step_count);
  const float c2 = 1.0f - powf(beta2, (float)step_count);
  const float g_norm = sqrtf(g_sq_norm);
#pragma omp parallel for schedule(static) if (E >= 200000)
  for (int64_t i = 0; i < E; ++i) {
    float scaled_grad = g[i] * rescale;
    scaled_grad /= g_norm;
    if (clip >= 0.0f) scaled_grad = clipf(scaled_grad, clip);
    float m = beta1 * mean[i] + (1.0f - beta1) * scaled_grad;
    float v = beta2 * var[i] + (1.0f - beta2) * scaled_grad * scaled_grad;
    mean[i] = m; var[i] = v;
    float mean_hat = m / c1;
    float var_hat = v / c2;
    var_hat = sqrtf(var_hat) + eps;
    float scaled_w = wd * w[i];
    temp_m[i] = mean_hat / var_hat + scaled_w;
    temp_g[i] = scaled_grad / var_hat + scaled_w;
  }
}
/* MultiLANSKernelStep2, multi_lans.cc:86-140 */
void kvo_lans_step2_f32(int64_t E, float* w, const float* temp_m, const float* temp_g, float lr, float beta1,
                        float sum_sq_w, float sum_sq_m, float sum_sq_g, float lower_bound, float upper_bound) {
  float r1 = sqrtf(sum_sq_w);
  float r2_m = sqrtf(sum_sq_m);
  float r2_g = sqrtf(sum_sq_g);
  if (lower_bound >= 0) r1 = r1 > lower_bound ? r1 : lower_bound;
  if (upper_bound >= 0) r1 = r1 < upper_bound ? r1 : upper_bound;
  float r_m = (r1 == 0.0f || r2_m == 0.0f) ? 1.0f : r1 / r2_m;
  float r_g = (r1 == 0.0f || r2_g == 0.0f) ? 1.0f : r1 / r2_g;
  r_m *= beta1;
  r_g *= (1. - beta1);                       /* double right-hand side, as written in the reference */
  float lr_adjusted_m = lr * r_m;
  float lr_adjusted_g = lr * r_g;
#pragma omp parallel for schedule(static) if (E >= 200000)
  for (int64_t i = 0; i < E; ++i) w[i] = w[i] - (lr_adjusted_m * temp_m[i] + lr_adjusted_g * temp_g[i]);
}

/* LARS._get_lars, lars.py:117-133: float32 NDArray arithmetic on the two norms; the caller multiplies
 * the (double) learning rate by the returned value */
float kvo_lars_ratio_f32(float sum_sq_w, float sum_sq_g, float eta, float wd, float eps) {
  float w_norm = sqrtf(sum_sq_w);
  float g_norm = sqrtf(sum_sq_g);
  float ratio = w_norm / g_norm;
  float lars = (eta * w_norm) / ((g_norm + wd * w_norm) + eps);
  if (!isfinite(ratio) || ratio == 0.0f) lars = 1.0f;    /* nan_or_zero = 1 - ratio / ratio */
  return lars;
}
