// Named arguments: sum(x, keywords::axis = -3), graph->param(name, shape,
// keywords::init = inits::zeros, keywords::fixed = true).
//
// Same call syntax as the reference's keywords (src/common/keywords.h), which
// is built on boost::any; this is a typed re-implementation: a key is a tag
// type, `key = value` yields a KeyVal, Get() scans the argument pack.
#pragma once

#include <functional>

#include "common/definitions.h"
#include "common/shape.h"

namespace marian {
namespace keywords {

template <class T, int Tag>
struct KeyVal {
  T value;
};

template <class T, int Tag>
struct Key {
  constexpr Key() {}
  template <class U>
  KeyVal<T, Tag> operator=(U&& v) const {
    return KeyVal<T, Tag>{T(std::forward<U>(v))};
  }
};

template <class T, int Tag>
T Get(Key<T, Tag>, T def) {
  return def;
}
template <class T, int Tag, class... Rest>
T Get(Key<T, Tag> k, T def, KeyVal<T, Tag> kv, Rest... rest) {
  return kv.value;
}
template <class T, int Tag, class First, class... Rest>
T Get(Key<T, Tag> k, T def, First, Rest... rest) {
  return Get(k, def, rest...);
}

template <class T, int Tag>
constexpr bool Has(Key<T, Tag>) {
  return false;
}
template <class T, int Tag, class... Rest>
constexpr bool Has(Key<T, Tag>, KeyVal<T, Tag>, Rest...) {
  return true;
}
template <class T, int Tag, class First, class... Rest>
constexpr bool Has(Key<T, Tag> k, First, Rest... rest) {
  return Has(k, rest...);
}

typedef KeyVal<int, 0> axis_k;
static const Key<int, 0> axis;
static const Key<std::function<void(Tensor)>, 1> init;
static const Key<bool, 2> fixed;
static const Key<Shape, 3> shape;
static const Key<Expr, 4> mask;
static const Key<float, 5> dropout_prob;
static const Key<float, 6> prob;
static const Key<bool, 7> final;

}  // namespace keywords
}  // namespace marian
