// ref_harness.cc -- links the REFERENCE's own arithmetic (header-only mshadow
// expression templates under /root/reference/3rdparty/mshadow) into a tiny C
// library, oracle/_ref/libkvref.so, so the oracle's dense sums can be pinned
// bit-for-bit against what libmxnet would compute.  TEST INFRASTRUCTURE ONLY.
//
// Nothing from the reference tree is copied: this file only #includes the
// headers where they lie (see oracle/Makefile) and spells the two call shapes:
//   * device order: ElementwiseSum<cpu>, src/ndarray/ndarray_function-inl.h:443-489
//     (mshadow expression `out = in0 + in1 (+ in2 (+ in3))`, else copy then `out += in_i`)
//   * CommCPU order: ReduceSumCPU, src/kvstore/comm.h:359-393
//     (`in_0 += in_1 + in_2 + in_3 + in_4` over groups of four)
// The build exists only in the authoring container (the GPU box has no
// /root/reference); the built .so travels with the snapshot.
#include <mshadow/tensor.h>
#include <cstdint>
#include <vector>

using mshadow::cpu;
using mshadow::Shape1;
using mshadow::Tensor;
using mshadow::index_t;

template <typename T>
static Tensor<cpu, 1, T> view(T* p, int64_t n) {
  return Tensor<cpu, 1, T>(p, Shape1(static_cast<index_t>(n)));
}

template <typename T>
static void sum_device_order(int n, T* const* src, int64_t E, T* out) {
  Tensor<cpu, 1, T> o = view(out, E);
  if (n == 1) {
    o = mshadow::expr::F<mshadow::op::identity>(view(src[0], E));
  } else if (n == 2) {
    o = view(src[0], E) + view(src[1], E);
  } else if (n == 3) {
    o = view(src[0], E) + view(src[1], E) + view(src[2], E);
  } else if (n == 4) {
    o = view(src[0], E) + view(src[1], E) + view(src[2], E) + view(src[3], E);
  } else {
    o = mshadow::expr::F<mshadow::op::identity>(view(src[0], E));
    for (int i = 1; i < n; ++i) o += view(src[i], E);
  }
}

template <typename T>
static void sum_commcpu_order(int n, T* const* p, int64_t off, int64_t size) {
  Tensor<cpu, 1, T> acc = view(p[0] + off, size);
  int i = 1;
  while (i < n) {
    int left = n - i;
    if (left >= 4) {
      acc += view(p[i] + off, size) + view(p[i + 1] + off, size) + view(p[i + 2] + off, size) +
             view(p[i + 3] + off, size);
    } else if (left == 3) {
      acc += view(p[i] + off, size) + view(p[i + 1] + off, size) + view(p[i + 2] + off, size);
    } else if (left == 2) {
      acc += view(p[i] + off, size) + view(p[i + 1] + off, size);
    } else {
      acc += view(p[i] + off, size);
    }
    i += 4;
  }
}

extern "C" {
void kvref_sum_device_f32(int n, float* const* src, int64_t E, float* out) {
  sum_device_order<float>(n, src, E, out);
}
void kvref_sum_device_f64(int n, double* const* src, int64_t E, double* out) {
  sum_device_order<double>(n, src, E, out);
}
void kvref_sum_device_i32(int n, int32_t* const* src, int64_t E, int32_t* out) {
  sum_device_order<int32_t>(n, src, E, out);
}
void kvref_sum_device_f16(int n, uint16_t* const* src, int64_t E, uint16_t* out) {
  typedef mshadow::half::half_t h;
  sum_device_order<h>(n, reinterpret_cast<h* const*>(src), E, reinterpret_cast<h*>(out));
}
// in place into p[0]; `threads`/`bound` as in comm.h:396-411
void kvref_sum_cpu_f32(int n, float* const* p, int64_t total, int threads, int64_t bound) {
  const int64_t step = bound < 4096 ? bound : 4096;
  const int64_t ntask = (total + step - 1) / step;
  if (total < bound || threads <= 1) {
    sum_commcpu_order<float>(n, p, 0, total);
  } else {
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t j = 0; j < ntask; ++j) {
      int64_t b = j * step < total ? j * step : total;
      int64_t e = (j + 1) * step < total ? (j + 1) * step : total;
      sum_commcpu_order<float>(n, p, b, e - b);
    }
  }
}
}
