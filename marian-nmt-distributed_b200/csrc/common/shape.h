// Host-side tensor shape (variable rank, numpy-style negative indexing).
// Same observable behaviour as the reference's marian::Shape
// (src/common/shape.h:19-192): default rank-1 {1}, dim(-1) = last, broadcast()
// right-aligns.  Kernels see the 4-D right-aligned POD view Shape4 below,
// which plays the role of gpu::ConstantShape<4> (src/gpu/shape.h:22-128).
#pragma once

#include <algorithm>
#include <initializer_list>
#include <string>
#include <vector>

#include "common/definitions.h"

#if defined(__CUDACC__)
#define MRN_HD __host__ __device__ __forceinline__
#else
#define MRN_HD inline
#endif

namespace marian {

struct Shape {
  std::vector<int> shape_;

  Shape() : shape_{1} {}
  Shape(std::initializer_list<int> il) : shape_(il) {}
  explicit Shape(const std::vector<int>& v) : shape_(v) {}

  void resize(size_t n) { shape_.resize(n, 1); }
  size_t size() const { return shape_.size(); }

  int& dim(int i) {
    int n = (int)size();
    int j = i >= 0 ? i : n + i;
    ABORT_IF(j < 0 || j >= n, "Shape index out of bounds:", i, "rank", n);
    return shape_[j];
  }
  const int& dim(int i) const { return const_cast<Shape&>(*this).dim(i); }
  int operator[](int i) const { return dim(i); }
  void set(int i, int v) { dim(i) = v; }
  int back() const { return shape_.back(); }

  int axis(int ax) const { return ax < 0 ? (int)size() + ax : ax; }

  int stride(int i) const {
    int j = axis(i);
    int s = 1;
    for(int k = (int)size() - 1; k > j; --k)
      s *= shape_[k];
    return s;
  }

  int elements() const {
    long el = 1;
    for(int s : shape_)
      el *= s;
    ABORT_IF(el > 2147483647L, "Tensor has more than 2^31-1 elements");
    return (int)el;
  }

  std::vector<int>::iterator begin() { return shape_.begin(); }
  std::vector<int>::iterator end() { return shape_.end(); }
  std::vector<int>::const_iterator begin() const { return shape_.begin(); }
  std::vector<int>::const_iterator end() const { return shape_.end(); }

  bool operator==(const Shape& o) const { return shape_ == o.shape_; }
  bool operator!=(const Shape& o) const { return !(*this == o); }

  std::string toString() const {
    std::string s = "shape=";
    for(size_t i = 0; i < size(); ++i)
      s += (i ? "x" : "") + std::to_string(shape_[i]);
    return s + " size=" + std::to_string(elements());
  }

  static Shape broadcast(const std::vector<Shape>& shapes) {
    size_t maxDims = 0;
    for(auto& s : shapes)
      maxDims = std::max(maxDims, s.size());
    Shape out;
    out.resize(maxDims);
    for(auto& s : shapes)
      for(int i = 1; i <= (int)s.size(); ++i) {
        ABORT_IF(out[-i] != s[-i] && out[-i] != 1 && s[-i] != 1,
                 "Shapes cannot be broadcasted:", out.toString(), s.toString());
        out.set(-i, std::max(out[-i], s[-i]));
      }
    return out;
  }

  // Works for anything with ->shape() (Tensor, Expr).
  template <class T>
  static Shape broadcast(const std::vector<T>& nodes) {
    std::vector<Shape> shapes;
    for(auto& n : nodes)
      shapes.push_back(n->shape());
    return broadcast(shapes);
  }
  template <class T>
  static Shape broadcast(std::initializer_list<T> il) {
    return broadcast(std::vector<T>(il));
  }
};

// 4-D right-aligned POD shape passed by value into kernels.
struct Shape4 {
  int d[4];
  int st[4];   // contiguous strides
  int bst[4];  // broadcast strides: 0 where the dim is 1

  MRN_HD Shape4() {
    for(int i = 0; i < 4; ++i) {
      d[i] = 1;
      st[i] = 1;
      bst[i] = 0;
    }
  }

  explicit Shape4(const Shape& s) {
    int n = (int)s.size();
    ABORT_IF(n > 4, "Tensors of rank > 4 are not supported:", s.toString());
    for(int i = 0; i < 4; ++i)
      d[i] = 1;
    for(int i = 0; i < n; ++i)
      d[4 - n + i] = s.shape_[i];
    update();
  }

  MRN_HD void update() {
    st[3] = 1;
    for(int i = 2; i >= 0; --i)
      st[i] = st[i + 1] * d[i + 1];
    for(int i = 0; i < 4; ++i)
      bst[i] = d[i] == 1 ? 0 : st[i];
  }

  MRN_HD int elements() const { return d[0] * d[1] * d[2] * d[3]; }
  MRN_HD int back() const { return d[3]; }

  MRN_HD void dims(int i, int* o) const {
    o[3] = i % d[3];
    i /= d[3];
    o[2] = i % d[2];
    i /= d[2];
    o[1] = i % d[1];
    o[0] = i / d[1];
  }
  MRN_HD int index(const int* o) const { return o[0] * st[0] + o[1] * st[1] + o[2] * st[2] + o[3] * st[3]; }
  MRN_HD int bindex(const int* o) const { return o[0] * bst[0] + o[1] * bst[1] + o[2] * bst[2] + o[3] * bst[3]; }

  MRN_HD bool operator==(const Shape4& o) const {
    return d[0] == o.d[0] && d[1] == o.d[1] && d[2] == o.d[2] && d[3] == o.d[3];
  }
  MRN_HD bool operator!=(const Shape4& o) const { return !(*this == o); }
};

}  // namespace marian
