// Element-wise functor DSL:  Element(_1 = _2 + _3, out, a, b),
// Add(_1 * (1.f - _2 * _2), grad, adj, val), ...
//
// Keeps the call syntax of the reference's functional:: namespace
// (src/functional/functional.h:12-20, operands.h:29-70, predicates.h:46-239)
// because node operators, optimizers and graph groups are written in it.  The
// implementation is different: an expression is a PURE function of a small
// array of operand values (`v[0]` = _1, `v[1]` = _2, ...).  The kernels load
// the operands (vectorised, with broadcast strides resolved per operand) into
// registers and call the expression once per element; `_1 = expr` simply
// yields expr, the kernel stores it.  Compile-time traits tell the kernels
// which operands are actually read, so `_1 = _2 + _3` never loads `out`.
#pragma once

#include <cmath>

#include "common/shape.h"

namespace marian {
namespace functional {

struct ExprTag {};

template <class T>
struct is_expr {
  static constexpr bool value = std::is_base_of<ExprTag, T>::value;
};

template <class X, class Y>
struct Assign : ExprTag {
  X x;
  Y y;
  Assign(X x_, Y y_) : x(x_), y(y_) {}
  MRN_HD float operator()(const float* v) const { return y(v); }
};

struct Capture : ExprTag {
  float value;
  Capture(float v) : value(v) {}
  MRN_HD float operator()(const float*) const { return value; }
};

template <class Op, class X>
struct Unary : ExprTag {
  X x;
  Unary(X x_) : x(x_) {}
  MRN_HD float operator()(const float* v) const { return Op::apply(x(v)); }
};

template <class Op, class X, class Y>
struct Binary : ExprTag {
  X x;
  Y y;
  Binary(X x_, Y y_) : x(x_), y(y_) {}
  MRN_HD float operator()(const float* v) const { return Op::apply(x(v), y(v)); }
};

namespace op {
// Two-branch sigmoid, as the reference's `logit` functor and stableLogit()
// (src/functional/predicates.h:92, src/kernels/tensor_operators.cu:15-23).
MRN_HD float sigmoid(float x) {
#if defined(__CUDA_ARCH__)
  // one exponential of a non-positive argument, branch free; ex2.approx / rcp.approx as the
  // reference's release build (--use_fast_math, CMakeLists.txt:37) compiles its functor
  float z = __expf(-fabsf(x));
  float r = __fdividef(1.f, 1.f + z);
  return x > 0.f ? r : z * r;
#else
  if(x > 0.f)
    return 1.f / (1.f + expf(-x));
  float z = expf(x);
  return z / (1.f + z);
#endif
}
struct Plus { MRN_HD static float apply(float a, float b) { return a + b; } };
struct Minus { MRN_HD static float apply(float a, float b) { return a - b; } };
struct Mult { MRN_HD static float apply(float a, float b) { return a * b; } };
struct Div { MRN_HD static float apply(float a, float b) { return a / b; } };
struct Neg { MRN_HD static float apply(float a) { return -a; } };
struct Tanh { MRN_HD static float apply(float a) { return tanhf(a); } };
struct Log { MRN_HD static float apply(float a) { return logf(a); } };
struct Exp { MRN_HD static float apply(float a) { return expf(a); } };
struct Sqrt { MRN_HD static float apply(float a) { return sqrtf(a); } };
struct Abs { MRN_HD static float apply(float a) { return fabsf(a); } };
struct Logit { MRN_HD static float apply(float a) { return sigmoid(a); } };
struct Sgn { MRN_HD static float apply(float a) { return (float)((0.f < a) - (a < 0.f)); } };
struct ReLU { MRN_HD static float apply(float a) { return a > 0.f ? a : 0.f; } };
struct ReLUback { MRN_HD static float apply(float a) { return a > 0.f ? 1.f : 0.f; } };
struct PReLU { MRN_HD static float apply(float a, float b) { return a > 0.f ? a : a * b; } };
struct PReLUback { MRN_HD static float apply(float a, float b) { return a > 0.f ? 1.f : b; } };
struct Pow { MRN_HD static float apply(float a, float b) { return powf(a, b); } };
// reference: BINARY(Clip, clip, fabs(x) >= y ? sgn(x) * y : x)  (predicates.h:118)
struct Clip {
  MRN_HD static float apply(float a, float b) { return fabsf(a) >= b ? Sgn::apply(a) * b : a; }
};
struct Gt { MRN_HD static float apply(float a, float b) { return a > b; } };
struct Lt { MRN_HD static float apply(float a, float b) { return a < b; } };
struct Geq { MRN_HD static float apply(float a, float b) { return a >= b; } };
struct Leq { MRN_HD static float apply(float a, float b) { return a <= b; } };
struct Eq { MRN_HD static float apply(float a, float b) { return a == b; } };
struct NEq { MRN_HD static float apply(float a, float b) { return a != b; } };
}  // namespace op

namespace detail {
template <class X, class = typename std::enable_if<is_expr<X>::value>::type>
X wrap(X x) {
  return x;
}
inline Capture wrap(float x) {
  return Capture(x);
}
}  // namespace detail

template <int N>
struct Var : ExprTag {
  MRN_HD Var() {}
  MRN_HD float operator()(const float* v) const { return v[N - 1]; }

  template <class X, class = typename std::enable_if<is_expr<X>::value>::type>
  Assign<Var<N>, X> operator=(X x) const {
    return Assign<Var<N>, X>(*this, x);
  }
  Assign<Var<N>, Capture> operator=(float x) const { return Assign<Var<N>, Capture>(*this, Capture(x)); }

  template <class X>
  auto operator+=(X x) const -> Assign<Var<N>, Binary<op::Plus, Var<N>, decltype(detail::wrap(x))>> {
    return Assign<Var<N>, Binary<op::Plus, Var<N>, decltype(detail::wrap(x))>>(*this, {*this, detail::wrap(x)});
  }
  template <class X>
  auto operator-=(X x) const -> Assign<Var<N>, Binary<op::Minus, Var<N>, decltype(detail::wrap(x))>> {
    return Assign<Var<N>, Binary<op::Minus, Var<N>, decltype(detail::wrap(x))>>(*this, {*this, detail::wrap(x)});
  }
  template <class X>
  auto operator*=(X x) const -> Assign<Var<N>, Binary<op::Mult, Var<N>, decltype(detail::wrap(x))>> {
    return Assign<Var<N>, Binary<op::Mult, Var<N>, decltype(detail::wrap(x))>>(*this, {*this, detail::wrap(x)});
  }
  template <class X>
  auto operator/=(X x) const -> Assign<Var<N>, Binary<op::Div, Var<N>, decltype(detail::wrap(x))>> {
    return Assign<Var<N>, Binary<op::Div, Var<N>, decltype(detail::wrap(x))>>(*this, {*this, detail::wrap(x)});
  }
};

static const Var<1> _1;
static const Var<2> _2;
static const Var<3> _3;
static const Var<4> _4;
static const Var<5> _5;

// ---- operator / function builders -------------------------------------
#define MRN_BINARY(OPNAME, FNAME)                                                             \
  template <class X, class Y,                                                                \
            class = typename std::enable_if<is_expr<X>::value && is_expr<Y>::value>::type>   \
  Binary<op::OPNAME, X, Y> FNAME(X x, Y y) {                                                 \
    return Binary<op::OPNAME, X, Y>(x, y);                                                   \
  }                                                                                          \
  template <class X, class = typename std::enable_if<is_expr<X>::value>::type>               \
  Binary<op::OPNAME, X, Capture> FNAME(X x, float y) {                                       \
    return Binary<op::OPNAME, X, Capture>(x, Capture(y));                                    \
  }                                                                                          \
  template <class Y, class = typename std::enable_if<is_expr<Y>::value>::type>               \
  Binary<op::OPNAME, Capture, Y> FNAME(float x, Y y) {                                       \
    return Binary<op::OPNAME, Capture, Y>(Capture(x), y);                                    \
  }

#define MRN_UNARY(OPNAME, FNAME)                                               \
  template <class X, class = typename std::enable_if<is_expr<X>::value>::type> \
  Unary<op::OPNAME, X> FNAME(X x) {                                            \
    return Unary<op::OPNAME, X>(x);                                            \
  }

MRN_BINARY(Plus, operator+)
MRN_BINARY(Minus, operator-)
MRN_BINARY(Mult, operator*)
MRN_BINARY(Div, operator/)
MRN_BINARY(Gt, operator>)
MRN_BINARY(Lt, operator<)
MRN_BINARY(Geq, operator>=)
MRN_BINARY(Leq, operator<=)
MRN_BINARY(Eq, operator==)
MRN_BINARY(NEq, operator!=)
MRN_BINARY(Pow, pow)
MRN_BINARY(Clip, clip)
MRN_BINARY(PReLU, PReLU)
MRN_BINARY(PReLUback, PReLUback)
MRN_UNARY(Neg, operator-)
MRN_UNARY(Tanh, tanh)
MRN_UNARY(Log, log)
MRN_UNARY(Exp, exp)
MRN_UNARY(Sqrt, sqrt)
MRN_UNARY(Abs, abs)
MRN_UNARY(Logit, logit)
MRN_UNARY(Sgn, sgn)
MRN_UNARY(ReLU, ReLU)
MRN_UNARY(ReLUback, ReLUback)

#undef MRN_BINARY
#undef MRN_UNARY

// ---- traits: does expression E read operand N (1-based)? ---------------
template <class E, int N>
struct Reads {
  static constexpr bool value = false;
};
template <int M, int N>
struct Reads<Var<M>, N> {
  static constexpr bool value = (M == N);
};
template <class Op, class X, int N>
struct Reads<Unary<Op, X>, N> {
  static constexpr bool value = Reads<X, N>::value;
};
template <class Op, class X, class Y, int N>
struct Reads<Binary<Op, X, Y>, N> {
  static constexpr bool value = Reads<X, N>::value || Reads<Y, N>::value;
};
template <class X, class Y, int N>
struct Reads<Assign<X, Y>, N> {
  static constexpr bool value = Reads<Y, N>::value;  // the assignee itself is only written
};

}  // namespace functional
}  // namespace marian
