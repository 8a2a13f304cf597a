// optim_math.h -- the per-element arithmetic of the fused optimizers, in a header that compiles both for
// the device (nvcc: the kernels in kernels.cu / rsp_kernels.cu / norm_kernels.cu) and for the host (g++:
// tests/c/optim_host.cc, which the CPU test-suite checks bit-for-bit against the oracle -- the same source
// the GPU executes, verified without a GPU).
//
// On the device the __f*_rn intrinsics keep ptxas from contracting a*b+c into an FMA.  On the host they
// are plain IEEE single-precision operations; the host translation unit MUST be built with
// -ffp-contract=off for the same reason.
#pragma once
#include "kernels.h"

#if defined(__CUDACC__)
#define MXKV_HD __device__ __forceinline__
#else
#include <cmath>
#define MXKV_HD inline
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
#endif

namespace mxkv {

struct Hyper {
  float lr, wd, eta, rescale, clip, momentum, beta1, beta2, eps;
};

MXKV_HD float clipf(float x, float b) {   // mshadow_op::clip, mshadow_op.h:999-1009
  return x > b ? b : (x < -b ? -b : x);
}

// one element of the fused update; returns the new weight.  Operation order is
// the reference source's, see the OptKind comments in kernels.h.
template <int OPT>
MXKV_HD float update_one(float g, float w, float& s0, float& s1, const Hyper& h) {
  if (OPT == OPT_SGD) {
    float r = __fmul_rn(h.rescale, g);
    if (h.clip >= 0.0f) r = clipf(r, h.clip);
    r = __fadd_rn(r, __fmul_rn(h.wd, w));
    return __fsub_rn(w, __fmul_rn(h.lr, r));
  } else if (OPT == OPT_SGD_MOM) {
    float r = __fmul_rn(h.rescale, g);
    if (h.clip >= 0.0f) r = clipf(r, h.clip);
    r = __fadd_rn(r, __fmul_rn(h.wd, w));
    float m = __fmul_rn(s0, h.momentum);
    m = __fsub_rn(m, __fmul_rn(h.lr, r));
    s0 = m;
    return __fadd_rn(w, m);
  } else if (OPT == OPT_ADAM) {
    float r = __fmul_rn(g, h.rescale);
    if (h.clip >= 0.0f) r = clipf(r, h.clip);
    r = __fadd_rn(r, __fmul_rn(w, h.wd));
    const float m = __fadd_rn(__fmul_rn(h.beta1, s0), __fmul_rn(__fsub_rn(1.f, h.beta1), r));
    const float v = __fadd_rn(__fmul_rn(h.beta2, s1),
                              __fmul_rn(__fmul_rn(__fsub_rn(1.f, h.beta2), r), r));
    s0 = m; s1 = v;
    return __fsub_rn(w, __fdiv_rn(__fmul_rn(h.lr, m), __fadd_rn(__fsqrt_rn(v), h.eps)));
  } else if (OPT == OPT_ADAMW) {
    float sg = __fmul_rn(h.rescale, g);
    if (h.clip >= 0.0f) sg = clipf(sg, h.clip);
    const float m = __fadd_rn(__fmul_rn(h.beta1, s0), __fmul_rn(__fsub_rn(1.0f, h.beta1), sg));
    const float v = __fadd_rn(__fmul_rn(h.beta2, s1),
                              __fmul_rn(__fsub_rn(1.0f, h.beta2), __fmul_rn(sg, sg)));
    s0 = m; s1 = v;
    const float step = __fadd_rn(__fdiv_rn(__fmul_rn(h.lr, m), __fadd_rn(__fsqrt_rn(v), h.eps)),
                                 __fmul_rn(h.wd, w));
    return __fsub_rn(w, __fmul_rn(h.eta, step));
  } else if (OPT == OPT_TEST) {
    const float gr = __fmul_rn(h.rescale, g);
    const float s = __fadd_rn(gr, __fmul_rn(h.wd, w));
    return __fsub_rn(w, __fmul_rn(h.lr, s));
  } else if (OPT == OPT_SGD_STD) {
    // every row: w *= (1 - lr*wd); rows of the gradient: w -= lr * clip(rescale*g) (wd already applied)
    const float ws = __fmul_rn(w, __fsub_rn(1.0f, __fmul_rn(h.lr, h.wd)));
    float r = __fmul_rn(h.rescale, g);
    if (h.clip >= 0.0f) r = clipf(r, h.clip);
    r = __fadd_rn(r, __fmul_rn(0.0f, ws));
    return __fsub_rn(ws, __fmul_rn(h.lr, r));
  } else if (OPT == OPT_ADAM_STD) {
    float r = __fmul_rn(g, h.rescale);
    if (h.clip >= 0.0f) r = clipf(r, h.clip);
    r = __fadd_rn(r, __fmul_rn(w, h.wd));
    const float m = __fadd_rn(__fmul_rn(h.beta1, s0), __fmul_rn(__fsub_rn(1.f, h.beta1), r));
    const float v = __fadd_rn(__fmul_rn(h.beta2, s1), __fmul_rn(__fsub_rn(1.f, h.beta2), __fmul_rn(r, r)));
    s0 = m; s1 = v;
    return __fsub_rn(w, __fdiv_rn(__fmul_rn(h.lr, m), __fadd_rn(__fsqrt_rn(v), h.eps)));
  }
  return g;  // OPT_NONE
}

// ---------------------------------------------------------------------------
// 1-bit / 2-bit gradient compression with error feedback (src/kvstore/gradient_compression-inl.h:44-227).
// Bit layout is the reference's: byte j of the stream holds values 4j..4j+3 (2-bit) or 8j..8j+7 (1-bit),
// first value in the most significant bits.
// ---------------------------------------------------------------------------
// code word w of the stream: residual += grad for its values, emit the codes, keep the quantisation error
template <int BITS>
MXKV_HD uint32_t quantize_word(const float* grad, float* residual, int64_t n, float thr, int64_t w) {
  constexpr int PER_WORD = 32 / BITS;
  uint32_t word = 0;
  const int64_t base = w * PER_WORD;
#pragma unroll
  for (int j = 0; j < PER_WORD; ++j) {
    const int64_t i = base + j;
    if (i >= n) break;
    float r = __fadd_rn(residual[i], grad[i]);
    const int byte = j / (8 / BITS);           // byte within the word (little endian in memory)
    const int slot = j % (8 / BITS);           // value within the byte, MSB first
    if (BITS == 2) {
      if (r >= thr) { word |= (0x3u << (6 - 2 * slot)) << (8 * byte); r = __fsub_rn(r, thr); }
      else if (r <= -thr) { word |= (0x2u << (6 - 2 * slot)) << (8 * byte); r = __fsub_rn(r, -thr); }
    } else {
      if (r > thr) { word |= (0x1u << (7 - slot)) << (8 * byte); r = __fsub_rn(r, 1.0f); }
      else r = __fadd_rn(r, 1.0f);
    }
    residual[i] = r;
  }
  return word;
}

// value i of a code stream
template <int BITS>
MXKV_HD float dequantize_value(const uint32_t* in, float thr, int64_t i) {
  constexpr int PER_WORD = 32 / BITS;
  const uint32_t word = in[i / PER_WORD];
  const int j = static_cast<int>(i % PER_WORD);
  const int byte = j / (8 / BITS), slot = j % (8 / BITS);
  const uint32_t b = (word >> (8 * byte)) & 0xffu;
  if (BITS == 2) {
    const uint32_t code = (b >> (6 - 2 * slot)) & 0x3u;
    return code == 0x3u ? thr : (code == 0x2u ? -thr : 0.0f);
  }
  return ((b >> (7 - slot)) & 0x1u) ? 1.0f : -1.0f;
}

}  // namespace mxkv
