// optim_host.cc -- the kernels' arithmetic headers (csrc/optim_math.h, csrc/norm_math.h) compiled for the
// HOST: the same source text the GPU executes, with the __f*_rn intrinsics mapped to plain IEEE operations
// (build with -ffp-contract=off).  tests/test_kernel_math_host.py loads this as a shared library and checks
// every optimizer flavour bit-for-bit against the oracle -- arithmetic parity without a GPU.
#include <cstdint>
#include "optim_math.h"
#include "norm_math.h"

using namespace mxkv;

template <int OPT>
static void run(int64_t n, const float* g, float* w, float* s0, float* s1, const Hyper& h) {
  float d0 = 0.f, d1 = 0.f;
  for (int64_t i = 0; i < n; ++i) w[i] = update_one<OPT>(g[i], w[i], s0 ? s0[i] : d0, s1 ? s1[i] : d1, h);
}

extern "C" {

int host_update(int opt, int64_t n, const float* g, float* w, float* s0, float* s1, float lr, float wd, float eta,
                float rescale, float clip, float momentum, float beta1, float beta2, float eps) {
  Hyper h;
  h.lr = lr; h.wd = wd; h.eta = eta; h.rescale = rescale; h.clip = clip; h.momentum = momentum;
  h.beta1 = beta1; h.beta2 = beta2; h.eps = eps;
  switch (opt) {
    case OPT_NONE: run<OPT_NONE>(n, g, w, s0, s1, h); break;
    case OPT_SGD: run<OPT_SGD>(n, g, w, s0, s1, h); break;
    case OPT_SGD_MOM: run<OPT_SGD_MOM>(n, g, w, s0, s1, h); break;
    case OPT_ADAM: run<OPT_ADAM>(n, g, w, s0, s1, h); break;
    case OPT_ADAMW: run<OPT_ADAMW>(n, g, w, s0, s1, h); break;
    case OPT_TEST: run<OPT_TEST>(n, g, w, s0, s1, h); break;
    case OPT_SGD_STD: run<OPT_SGD_STD>(n, g, w, s0, s1, h); break;
    case OPT_ADAM_STD: run<OPT_ADAM_STD>(n, g, w, s0, s1, h); break;
    default: return -1;
  }
  return 0;
}

static void fill(NormLaunch* L, NormWork* tw, float lr, double lr_d, float wd, float c1, float c2, float rescale,
                 float clip, float beta1, float beta2, float eps, float lower, float upper, float lars_eta,
                 float lars_eps, int bias_correction, int flags, const float* totals) {
  *L = NormLaunch();
  *tw = NormWork();
  L->rescale = rescale; L->clip = clip; L->beta1 = beta1; L->beta2 = beta2; L->eps = eps;
  L->lower_bound = lower; L->upper_bound = upper; L->lars_eta = lars_eta; L->lars_eps = lars_eps;
  L->bias_correction = bias_correction;
  tw->lr = lr; tw->lr_d = lr_d; tw->wd = wd; tw->c1 = c1; tw->c2 = c2; tw->flags = flags;
  tw->norm_world = 1;
  tw->nrm_peer[0] = totals;
}

// LAMB step 1 (kv_norm_first_kernel / kv_norm_mid_kernel<NORM_LAMB>): update direction + moments
void host_lamb_step1(int64_t n, const float* g, const float* w, float* mean, float* var, float* ghat, float wd,
                     float c1, float c2, float rescale, float clip, float beta1, float beta2, float eps,
                     int bias_correction) {
  NormLaunch L; NormWork tw;
  fill(&L, &tw, 0.f, 0.0, wd, c1, c2, rescale, clip, beta1, beta2, eps, -1.f, -1.f, 0.f, 0.f, bias_correction, 0, nullptr);
  for (int64_t i = 0; i < n; ++i) ghat[i] = lamb_step1(g[i], w[i], mean[i], var[i], L, tw);
}

// LANS step 1 (kv_norm_mid_kernel<NORM_LANS>)
void host_lans_step1(int64_t n, const float* g, const float* w, float* mean, float* var, float* temp_m,
                     float* temp_g, float g_sq_norm, float wd, float c1, float c2, float rescale, float clip,
                     float beta1, float beta2, float eps) {
  NormLaunch L; NormWork tw;
  fill(&L, &tw, 0.f, 0.0, wd, c1, c2, rescale, clip, beta1, beta2, eps, -1.f, -1.f, 0.f, 0.f, 1, 0, nullptr);
  const float g_norm = __fsqrt_rn(g_sq_norm);
  for (int64_t i = 0; i < n; ++i) lans_step1(g[i], w[i], mean[i], var[i], g_norm, L, tw, temp_m[i], temp_g[i]);
}

// the per-key scalars of kv_norm_apply_kernel: flavor 0 LAMB (sc[0] = lr * ratio), 1 LANS (sc[0], sc[1]),
// 2 LARS (sc[0] = effective learning rate).  totals: the kNrm* slots.
void host_apply_scalars(int flavor, const float* totals, float lr, double lr_d, float wd, float lower, float upper,
                        float beta1, float lars_eta, float lars_eps, int flags, float* sc) {
  NormLaunch L; NormWork tw;
  fill(&L, &tw, lr, lr_d, wd, 1.f, 1.f, 1.f, -1.f, beta1, 0.f, 0.f, lower, upper, lars_eta, lars_eps, 1, flags, totals);
  sc[0] = sc[1] = 0.f;
  if (flavor == 0) apply_scalars<APPLY_LAMB>(tw, L, sc);
  else if (flavor == 1) apply_scalars<APPLY_LANS>(tw, L, sc);
  else apply_scalars<APPLY_LARS>(tw, L, sc);
}

int host_not_finite(float x) { return not_finite(x) ? 1 : 0; }

}  // extern "C"
