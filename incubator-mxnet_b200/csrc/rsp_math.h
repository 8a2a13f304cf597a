// rsp_math.h -- per-element arithmetic of the lazy row-wise updates (host/device, like optim_math.h).
#pragma once
#include "rsp_kernels.h"
#include "optim_math.h"

namespace mxkv {

// SGDDnsRspKernel / SGDMomDnsRspDnsKernel / AdamDnsRspDnsKernel (optimizer_op-inl.h:414-465,1305-1360): the
// arithmetic of the dense kernels, applied only to the rows present in the gradient.  `acc` is the merged
// gradient element, `wv` the current weight, `off` its offset in the table (and in the state arrays);
// returns the new weight.
template <int OPT>
MXKV_HD float rsp_lazy_update(float acc, float wv, int64_t off, const RspRowArgs& A) {
  float g = (OPT == OPT_ADAM) ? __fmul_rn(acc, A.rescale) : __fmul_rn(A.rescale, acc);
  if (A.clip >= 0.0f) g = clipf(g, A.clip);
  g = __fadd_rn(g, (OPT == OPT_ADAM) ? __fmul_rn(wv, A.wd) : __fmul_rn(A.wd, wv));
  if (OPT == OPT_SGD) {
    return __fsub_rn(wv, __fmul_rn(A.lr, g));
  } else if (OPT == OPT_SGD_MOM) {
    float* mp = A.s0 + off;
    float m = __fmul_rn(*mp, A.momentum);
    m = __fsub_rn(m, __fmul_rn(A.lr, g));
    *mp = m;
    return __fadd_rn(wv, m);
  } else {   // OPT_ADAM
    float* mp = A.s0 + off;
    float* vp = A.s1 + off;
    const float m = __fadd_rn(__fmul_rn(A.beta1, *mp), __fmul_rn(__fsub_rn(1.f, A.beta1), g));
    const float v = __fadd_rn(__fmul_rn(A.beta2, *vp), __fmul_rn(__fmul_rn(__fsub_rn(1.f, A.beta2), g), g));
    *mp = m; *vp = v;
    return __fsub_rn(wv, __fdiv_rn(__fmul_rn(A.lr, m), __fadd_rn(__fsqrt_rn(v), A.eps)));
  }
}

}  // namespace mxkv
