// Basic typedefs shared by the whole engine.
//
// Mirrors the vocabulary of the reference (src/common/definitions.h:38-60:
// Ptr/New/Weak, Tensor, Expr) so that model code written against Marian's
// operator API reads the same here.  Implementation is independent.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace marian {

template <class T>
using Ptr = std::shared_ptr<T>;
template <class T>
using Weak = std::weak_ptr<T>;

template <class T, typename... Args>
Ptr<T> New(Args&&... args) {
  return Ptr<T>(new T(std::forward<Args>(args)...));
}
template <class T>
Ptr<T> New(Ptr<T> p) {
  return Ptr<T>(p);
}

class TensorBase;
typedef Ptr<TensorBase> Tensor;

template <class DataType>
struct Chainable;
typedef Ptr<Chainable<Tensor>> Expr;
typedef Weak<Chainable<Tensor>> WExpr;

// Reference: NEMATUS_LN_EPS in src/common/definitions.h:95
const float NEMATUS_LN_EPS = 1e-5f;

// Errors.  The reference aborts the process (src/common/logging.h:43-65); a
// library behind a C ABI must not, so we throw and the C ABI translates the
// exception into a status code + message (include/marian_b200.h).
struct MarianError : public std::runtime_error {
  explicit MarianError(const std::string& m) : std::runtime_error(m) {}
};

namespace detail {
inline void fmt_into(std::ostringstream&) {}
template <class T, class... R>
void fmt_into(std::ostringstream& os, const T& t, const R&... r) {
  os << " " << t;
  fmt_into(os, r...);
}
}  // namespace detail

template <class... Args>
[[noreturn]] void abort_with(const char* file, int line, const std::string& msg, const Args&... args) {
  std::ostringstream os;
  os << msg;
  detail::fmt_into(os, args...);
  os << " [" << file << ":" << line << "]";
  throw MarianError(os.str());
}

#define ABORT(...) ::marian::abort_with(__FILE__, __LINE__, __VA_ARGS__)
#define ABORT_IF(cond, ...)                                \
  do {                                                     \
    if(cond)                                               \
      ::marian::abort_with(__FILE__, __LINE__, __VA_ARGS__); \
  } while(0)

}  // namespace marian
