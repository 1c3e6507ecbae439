// TEST ORACLE — not part of the product.  CPU restatement of the reference's
// Element / Add / Reduce templates (src/kernels/tensor_operators.h:26-274).
//
// This header SHADOWS csrc/kernels/element.h when the oracle library is built
// (oracle/Makefile puts -I oracle/cpu first), so the host graph code links
// against plain loops instead of CUDA kernels.
//
// Follows the reference's dispatch rules exactly:
//   Element: out[i] = f(out[i], in1[bindex(dims(i))], ...)     (gElement :26-49)
//   Add:  (1) full.back()!=1 && out.back()==1 -> last-axis reduction (gAddReduce :149-205)
//         (2) out.shape()==full              -> element-wise accumulate (gAddEqual :123-147)
//         (3) otherwise                      -> generic nested-loop reduction (gAddGeneric :90-121,
//                                               loop order of src/gpu/tmp.h:96-142)
//   Reduce = out->set(0); Add(...)                               (:259-274)
// Reductions accumulate in double and round once, so the oracle is the more
// accurate side of any comparison (the reference's own tree order in fp32 is
// not reproducible off-GPU anyway).
#pragma once

#include <vector>

#include "common/shape.h"
#include "functional/functional.h"
#include "tensors/tensor.h"

namespace marian {

namespace cpu {
struct View {
  float* p;
  Shape4 s;
};
inline View view(Tensor t) {
  return View{t->data(), Shape4(t->shape())};
}
}  // namespace cpu

template <class Functor, class... Tensors>
void Element(Functor functor, Tensor out, Tensors... tensors) {
  constexpr size_t K = sizeof...(tensors) + 1;
  cpu::View t[K] = {cpu::view(out), cpu::view(tensors)...};
  int length = t[0].s.elements();
  bool broadcast = false;
  for(size_t i = 1; i < K; ++i)
    broadcast = broadcast || t[0].s != t[i].s;

#pragma omp parallel for if(length > 16384)
  for(int index = 0; index < length; ++index) {
    float v[K];
    int dims[4];
    if(broadcast)
      t[0].s.dims(index, dims);
    v[0] = t[0].p[index];
    for(size_t i = 1; i < K; ++i)
      v[i] = t[i].p[broadcast ? t[i].s.bindex(dims) : index];
    t[0].p[index] = functor(v);
  }
}

template <class Functor, class... Tensors>
void Add(Functor functor, float scale, Tensor out, Tensors... tensors) {
  constexpr size_t K = sizeof...(Tensors);
  std::vector<Shape> shapes = {out->shape(), tensors->shape()...};
  Shape4 full(Shape::broadcast(shapes));
  cpu::View o = cpu::view(out);
  cpu::View in[K] = {cpu::view(tensors)...};
  int length = o.s.elements();

  if(full.back() != 1 && o.s.back() == 1) {
    // (1) reduce the last axis
    int rows = full.elements() / full.back();
    int cols = full.back();
#pragma omp parallel for if((long)rows * cols > 16384)
    for(int j = 0; j < rows; ++j) {
      double sum = 0;
      for(int id = 0; id < cols; ++id) {
        int dims[4];
        full.dims(j * cols + id, dims);
        float v[K];
        for(size_t i = 0; i < K; ++i)
          v[i] = in[i].p[in[i].s.bindex(dims)];
        sum += functor(v);
      }
      o.p[j] += (float)(sum * scale);
    }
  } else if(o.s == full) {
    // (2) same shape as the broadcast: accumulate element-wise
#pragma omp parallel for if(length > 16384)
    for(int index = 0; index < length; ++index) {
      int dims[4];
      o.s.dims(index, dims);
      float v[K];
      for(size_t i = 0; i < K; ++i)
        v[i] = in[i].p[in[i].s.bindex(dims)];
      o.p[index] += functor(v) * scale;
    }
  } else {
    // (3) generic: every output element sums over the reduced sub-space
    int len[4];
    for(int i = 0; i < 4; ++i)
      len[i] = full.d[i] / o.s.d[i];
#pragma omp parallel for if((long)full.elements() > 16384)
    for(int index = 0; index < length; ++index) {
      int od[4];
      o.s.dims(index, od);
      double sum = 0;
      int d[4];
      for(int i0 = 0; i0 < len[0]; ++i0)
        for(int i1 = 0; i1 < len[1]; ++i1)
          for(int i2 = 0; i2 < len[2]; ++i2)
            for(int i3 = 0; i3 < len[3]; ++i3) {
              d[0] = od[0] + i0;
              d[1] = od[1] + i1;
              d[2] = od[2] + i2;
              d[3] = od[3] + i3;
              float v[K];
              for(size_t i = 0; i < K; ++i)
                v[i] = in[i].p[in[i].s.bindex(d)];
              sum += functor(v);
            }
      o.p[index] += (float)(sum * scale);
    }
  }
}

template <class Functor, class... Tensors>
void Add(Functor functor, Tensor out, Tensors... tensors) {
  Add(functor, 1.f, out, tensors...);
}

template <class Functor, class... Tensors>
void Reduce(Functor functor, float scale, Tensor out, Tensors... tensors) {
  out->set(0);
  Add(functor, scale, out, tensors...);
}

template <class Functor, class... Tensors>
void Reduce(Functor functor, Tensor out, Tensors... tensors) {
  out->set(0);
  Add(functor, 1.f, out, tensors...);
}

}  // namespace marian
