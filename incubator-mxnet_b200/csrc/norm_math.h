// norm_math.h -- scalar arithmetic of the layer-wise adaptive optimizers (LAMB / LANS step 1 and 2, the
// LARS ratio, trust-ratio rules); like optim_math.h it compiles for the device and for the host so that the
// CPU test-suite can check the very source the kernels execute against the oracle.
#pragma once
#include "norm_kernels.h"
#include "optim_math.h"

#if !defined(__CUDACC__) && !defined(MXKV_HOST_EMU)     // (tests/sim/host_emu.h brings its own)
#include <cstring>
static inline unsigned int __float_as_uint(float x) { unsigned int u; std::memcpy(&u, &x, 4); return u; }
#endif

namespace mxkv {

MXKV_HD bool not_finite(float x) {
  return (__float_as_uint(x) & 0x7f800000u) == 0x7f800000u;
}

// mean/var update shared by LAMB and LANS step 1 (multi_lamb.cc:52-61, multi_lans.cc:58-66)
MXKV_HD void moments(float sg, float& mean, float& var, float beta1, float beta2) {
  mean = __fadd_rn(__fmul_rn(beta1, mean), __fmul_rn(__fsub_rn(1.0f, beta1), sg));
  var = __fadd_rn(__fmul_rn(beta2, var), __fmul_rn(__fmul_rn(__fsub_rn(1.0f, beta2), sg), sg));
}

// LAMB step 1 for one element (multi_lamb.cc:48-78); returns the update direction
MXKV_HD float lamb_step1(float g, float w, float& mean, float& var, const NormLaunch& L,
                                            const NormWork& tw) {
  float sg = __fmul_rn(g, L.rescale);
  if (L.clip >= 0.0f) sg = clipf(sg, L.clip);
  moments(sg, mean, var, L.beta1, L.beta2);
  if (L.bias_correction) {
    const float mean_hat = __fdiv_rn(mean, tw.c1);
    const float var_hat = __fdiv_rn(var, tw.c2);
    return __fadd_rn(__fdiv_rn(mean_hat, __fadd_rn(__fsqrt_rn(var_hat), L.eps)), __fmul_rn(tw.wd, w));
  }
  return __fadd_rn(__fdiv_rn(mean, __fadd_rn(__fsqrt_rn(var), L.eps)), __fmul_rn(tw.wd, w));
}

// LANS step 1 for one element (multi_lans.cc:48-84): the two update directions temp_m and temp_g
MXKV_HD void lans_step1(float g, float w, float& mean, float& var, float g_norm, const NormLaunch& L,
                        const NormWork& tw, float& temp_m, float& temp_g) {
  float sg = __fmul_rn(g, L.rescale);
  sg = __fdiv_rn(sg, g_norm);
  if (L.clip >= 0.0f) sg = clipf(sg, L.clip);
  moments(sg, mean, var, L.beta1, L.beta2);
  const float mean_hat = __fdiv_rn(mean, tw.c1);
  float var_hat = __fdiv_rn(var, tw.c2);
  var_hat = __fadd_rn(__fsqrt_rn(var_hat), L.eps);
  const float scaled_w = __fmul_rn(tw.wd, w);
  temp_m = __fadd_rn(__fdiv_rn(mean_hat, var_hat), scaled_w);
  temp_g = __fadd_rn(__fdiv_rn(sg, var_hat), scaled_w);
}

// total of one slot over the contributing ranks, in rank order (identical on every rank)
MXKV_HD float rank_total(const NormWork& tw, int slot) {
  float t = 0.f;
  for (int q = 0; q < tw.norm_world; ++q) t = __fadd_rn(t, tw.nrm_peer[q][slot]);
  return t;
}

enum ApplyFlavor : int { APPLY_LAMB = 0, APPLY_LANS = 1, APPLY_LARS = 2, APPLY_LARS_MOM = 3 };

// r1 with the optional bounds, then r1/r2 or 1 (multi_lamb.cc:96-109, multi_lans.cc:110-125)
MXKV_HD float bounded(float r1, const NormLaunch& L) {
  if (L.lower_bound >= 0.f) r1 = fmaxf(r1, L.lower_bound);
  if (L.upper_bound >= 0.f) r1 = fminf(r1, L.upper_bound);
  return r1;
}
MXKV_HD float trust(float r1, float r2) {
  return (r1 == 0.0f || r2 == 0.0f) ? 1.0f : __fdiv_rn(r1, r2);
}

template <int FLAVOR>
MXKV_HD void apply_scalars(const NormWork& tw, const NormLaunch& L, float* sc) {
  if (FLAVOR == APPLY_LAMB) {
    const float r1 = bounded(__fsqrt_rn(rank_total(tw, kNrmW)), L);
    const float r2 = __fsqrt_rn(rank_total(tw, kNrmG));
    sc[0] = __fmul_rn(tw.lr, trust(r1, r2));
  } else if (FLAVOR == APPLY_LANS) {
    const float r1 = bounded(__fsqrt_rn(rank_total(tw, kNrmW)), L);
    const float r2m = __fsqrt_rn(rank_total(tw, kNrmM));
    const float r2g = __fsqrt_rn(rank_total(tw, kNrmG2));
    float r_m = trust(r1, r2m);
    float r_g = trust(r1, r2g);
    r_m = __fmul_rn(r_m, L.beta1);
    // `r_g *= (1. - static_cast<MPDType>(beta1))`: the right-hand side is a double (multi_lans.cc:127)
    r_g = static_cast<float>(static_cast<double>(r_g) * (1.0 - static_cast<double>(L.beta1)));
    sc[0] = __fmul_rn(tw.lr, r_m);
    sc[1] = __fmul_rn(tw.lr, r_g);
  } else {   // LARS, lars.py:117-133: float32 NDArray arithmetic, then lr (a Python double) *= lars
    if (tw.flags & 2) {
      sc[0] = tw.lr;
    } else {
      const float w_norm = __fsqrt_rn(rank_total(tw, kNrmW));
      const float g_norm = __fsqrt_rn(rank_total(tw, kNrmG));
      const float ratio = __fdiv_rn(w_norm, g_norm);
      float lars = __fdiv_rn(__fmul_rn(L.lars_eta, w_norm),
                             __fadd_rn(__fadd_rn(g_norm, __fmul_rn(tw.wd, w_norm)), L.lars_eps));
      if (not_finite(ratio) || ratio == 0.0f) lars = 1.0f;    // nan_or_zero = 1 - ratio / ratio
      sc[0] = static_cast<float>(tw.lr_d * static_cast<double>(lars));
    }
  }
}

}  // namespace mxkv
